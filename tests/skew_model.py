"""
CPU model of the schedule of `mrtm_skew_kernel` (xanthos_b200/csrc/mrtm_skew.cu).  TEST INFRASTRUCTURE ONLY.

It executes, with numpy float64 arithmetic, exactly what one warp of the kernel does per loop iteration - driven by
the plan tables the C library exports (`xan_mrtm_skew_tables`): per-cell lags, the double-buffered flow table, ghost
imports / exports through per-edge series, the per-cell month events.  `tests/test_host.py` compares its output bit
for bit with `oracle.mrtm.route` (the restatement of xanthos/routing/mrtm.py:16-82), which pins the host-side plan
and the time-skew logic without a GPU; the CUDA kernel is then compared with the same oracle on the GPU.
"""

import ctypes

import numpy as np


def skew_tables(um):
    """Plan tables of an UpstreamMatrix (xanthos_b200.routing.mrtm.upstream_genmatrix) as a dict of int arrays."""
    from xanthos_b200 import _cuda as C
    lib = C.lib()
    info = (ctypes.c_int * 12)()
    C.check(lib.xan_mrtm_skew_info(um._plan, info))
    info = list(info)
    K, nw, ne = info[0], info[1], info[2]
    if nw == 0:
        return None
    ns, xg, xo = info[7], info[9], info[10]
    shapes = dict(cell=(nw, 32 * K), lag=(nw, 32 * K), src=(nw, 32, ns), ghost_edge=(nw, xg), ghost_lag=(nw, xg),
                  exp_edge=(nw, xo), exp_place=(nw, xo), Dw=(nw,), edge_prod=(max(ne, 1),), edge_cons=(max(ne, 1),))
    t = {k: np.zeros(v, dtype=np.int32) for k, v in shapes.items()}
    ptr = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))     # noqa: E731
    C.check(lib.xan_mrtm_skew_tables(um._plan, *[ptr(t[k]) for k in ('cell', 'lag', 'src', 'ghost_edge', 'ghost_lag',
                                                                      'exp_edge', 'exp_place', 'Dw', 'edge_prod',
                                                                      'edge_cons')]))
    t['edge_prod'], t['edge_cons'] = t['edge_prod'][:ne], t['edge_cons'][:ne]
    t['info'] = dict(K=K, n_warps=nw, n_edges=ne, n_levels=info[3], G=info[4], Dmax=info[5], n_pieces=info[6],
                     nsrc=ns, zero=info[8], XG=xg, XO=xo, lagm=info[11])
    return t


def route(t, q, flow_dist, velocity, area, ndays, dt, spinup, chs_prev=None):
    """q [ncell, nmonths] -> (ChStorage, Avg_ChFlow [ncell, nmonths], instream flow [ncell]) by the skewed schedule."""
    inf = t['info']
    K, nw, ne, ZERO, XG, LAGM = inf['K'], inf['n_warps'], inf['n_edges'], inf['zero'], inf['XG'], inf['lagm']
    CAP = 32 * K
    ncell, nmonths = q.shape
    steps = list(range(spinup)) + list(range(nmonths))                  # runoff month of every routing step
    nt = np.array([int(ndays[m] * 24 * 3600 / dt) for m in steps])
    secs = np.array([float(ndays[m] * 24 * 3600) for m in steps])
    start = np.concatenate([[0], np.cumsum(nt)])
    M, T = len(steps), int(start[-1])
    assert inf['Dmax'] < nt.min(), "event windows of consecutive months must not overlap"
    chs = np.zeros((ncell, nmonths))
    avg = np.zeros((ncell, nmonths))
    inst = np.zeros(ncell)
    ring = np.zeros((max(ne, 1), T, 2))
    dtinv = 1.0 / dt
    # producers before consumers
    indeg = np.zeros(nw, dtype=int)
    for e in range(ne):
        indeg[t['edge_cons'][e]] += 1
    order = [w for w in range(nw) if indeg[w] == 0]
    for w in order:
        for e in range(ne):
            if t['edge_prod'][e] == w:
                c = t['edge_cons'][e]
                indeg[c] -= 1
                if indeg[c] == 0:
                    order.append(c)
    assert len(order) == nw, "warp dependency graph has a cycle"

    lanes = np.arange(32)
    for w in order:
        cell, lag, src, Dw = t['cell'][w], t['lag'][w], t['src'][w], int(t['Dw'][w])
        has = cell >= 0
        cidx = np.where(has, cell, 0)
        S = np.zeros(CAP)
        ti = np.zeros(CAP)
        erl = np.zeros(CAP)
        fav = np.zeros(CAP)
        exch = np.zeros((LAGM + 1, CAP + XG + 1, 2))        # a reader reads what was written LAGM iterations earlier
        tauinv = np.where(has, velocity[cidx] / flow_dist[cidx], 0.0)
        b, evt = 0, int(start[0])
        for n in range(T + Dw + 1):
            rd, wr, last = exch[(n - LAGM) % (LAGM + 1)], exch[n % (LAGM + 1)], exch[(n - 1) % (LAGM + 1)]
            # ---- per-cell month events ------------------------------------------------------------------------
            if b <= M and n >= evt:
                k = n - evt
                hit = has & (lag == k)
                if hit.any():
                    c = cidx[hit]
                    if b > 0:
                        mprev = b - 1
                        if mprev >= spinup:
                            chs[c, mprev - spinup] = S[hit]
                            avg[c, mprev - spinup] = fav[hit] / nt[mprev]
                    if b == M:
                        inst[c] = last[np.nonzero(hit)[0], 1]
                    else:
                        fav[hit] = 0.0
                        erl[hit] = (q[c, steps[b]] * area[c]) * (1e6 / 1e3) / secs[b]
                    if b == 0:
                        S[hit] = chs_prev[c] if chs_prev is not None else 0.0
                        ti[hit] = tauinv[hit]
                if k == Dw:
                    b += 1
                    evt = int(start[b]) if b <= M else 1 << 60
            # ---- ghost imports and exports ---------------------------------------------------------------------
            for g in range(XG):
                e = t['ghost_edge'][w, g]
                if e >= 0:
                    tau = n - int(t['ghost_lag'][w, g])
                    wr[CAP + g] = ring[e, tau] if 0 <= tau < T else 0.0
            for o in range(t['exp_edge'].shape[1]):
                e = t['exp_edge'][w, o]
                tau = n - 1 - Dw
                if e >= 0 and 0 <= tau < T:
                    ring[e, tau] = last[t['exp_place'][w, o]]
            # ---- the K slots --------------------------------------------------------------------------------------
            F = S * ti
            d = np.empty(CAP)
            d2 = np.empty(CAP)
            x, y = rd[:, 0], rd[:, 1]
            for v, out in ((x, d), (y, d2)):
                s0 = src[:, :8]
                acc = v[s0[:, 0]]
                for j in (1, 2, 3):
                    acc = acc + v[s0[:, j]]
                acc = acc - F[:32]
                for j in (4, 5, 6, 7):
                    acc = acc + v[s0[:, j]]
                out[:32] = acc + erl[:32]
                for s in range(1, K):
                    sl = slice(32 * s, 32 * s + 32)
                    out[sl] = (v[src[:, 8 + s - 1]] - F[sl]) + erl[sl]
            with np.errstate(invalid='ignore', over='ignore'):
                clamp = (d * dt) < (-S)
                Sn = S + d2 * dt
                Fc = (d + F) + S * dtinv
            Fp = np.where(clamp, Fc, F)
            S = np.where(clamp, 0.0, Sn)
            fav = fav + Fp
            wr[:CAP, 0] = F
            wr[:CAP, 1] = Fp
    return chs, avg, inst


def pacing_window(um, requested, nt_min):
    """(effective window in months, sub-steps per hand-over chunk, ring entries, default request) of the library."""
    from xanthos_b200 import _cuda as C
    out = (ctypes.c_int * 4)()
    C.check(C.lib().xan_mrtm_skew_window(um._plan, int(requested), int(nt_min), out))
    return tuple(out)


def handover_completes(t, nt, window, ch, rl):
    """
    Replays the hand-over protocol of `mrtm_skew_kernel` (no arithmetic) on the plan tables `t`: does every warp reach
    the end of the run?  `nt`: sub-steps of every routing step (spin-up months, then months).  Per warp, as in the kernel:
      * before the loop a warp with ghost entries waits for `min(T, ch)` sub-steps of its producers;
      * at every chunk start n0 (multiples of `ch` up to T + Dw) a linked warp waits until its producers have handed over
        `min(T, n0 + 2 ch)` sub-steps and its consumers have consumed `n0 - rl` (ring back-pressure), then publishes
        `max(0, min(T, n0 - 1 - Dw))`;
      * in iteration `start[b] + Dw` it counts itself past month boundary b and, with a pacing window W > 0, waits until
        ALL warps are past boundary b - W;
      * after the loop it publishes T.
    Returns True when all warps finish, False on a deadlock (a full pass over the warps without any progress).
    """
    nw = t['info']['n_warps']
    Dw = t['Dw'].astype(int)
    prods = [[] for _ in range(nw)]
    conss = [[] for _ in range(nw)]
    for p, c in zip(t['edge_prod'], t['edge_cons']):
        prods[int(c)].append(int(p))
        conss[int(p)].append(int(c))
    start = np.concatenate([[0], np.cumsum(nt)]).astype(int)
    M, T = len(nt), int(start[-1])
    # per warp: the ordered list of blocking points (kind, a, b, publish)
    plans = []
    for w in range(nw):
        linked = bool(prods[w] or conss[w])
        steps = []
        if prods[w]:
            steps.append((-1, 'P', min(T, ch), 0, None))
        nlast = T + Dw[w]
        for n0 in range(0, nlast + 1, ch):
            if linked:
                steps.append((n0 - 0.5, 'P', min(T, n0 + 2 * ch), n0 - rl, max(0, min(T, n0 - 1 - Dw[w]))))
        for b in range(M + 1):
            steps.append((start[b] + Dw[w], 'B', b, 0, None))
        steps.sort(key=lambda s: s[0])
        if linked:
            steps.append((nlast + ch, 'F', 0, 0, T))
        plans.append(steps)
    progress = np.zeros(nw, dtype=int)
    done = np.zeros(M + 2, dtype=int)
    pc = [0] * nw
    counted = [set() for _ in range(nw)]
    live = set(range(nw))
    while live:
        moved = False
        for w in list(live):
            steps = plans[w]
            while pc[w] < len(steps):
                _, kind, a, b, pub = steps[pc[w]]
                if kind == 'P':
                    if any(progress[p] < a for p in prods[w]) or (b > 0 and any(progress[c] < b for c in conss[w])):
                        break
                    if pub is not None:
                        progress[w] = pub
                elif kind == 'B':
                    if a not in counted[w]:
                        counted[w].add(a)
                        done[a] += 1
                        moved = True
                    if window > 0 and a - window >= 0 and done[a - window] < nw:
                        break
                else:
                    progress[w] = pub
                pc[w] += 1
                moved = True
            if pc[w] == len(steps):
                live.discard(w)
        if not moved:
            return False
    return True
