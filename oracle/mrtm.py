"""
MRTM river-routing oracle (test infrastructure only).

Restates xanthos/routing/mrtm.py: `downstream` (:85-120) with
`make_flowdirgrid` (:233-258), `upstream` (:123-191), the sparse operator of
`upstream_genmatrix` (:194-230) as an explicit per-row gather list, the monthly
explicit-Euler step `streamrouting` (:16-82) and the spin-up + simulation
driver of Components.calculate_routing (xanthos/components.py:262-296).

scipy.sparse is not used: row i of UM = UP - I holds +1 at every upstream
cell of i and -1 at i itself, stored in ascending column order, and scipy's
CSR matvec accumulates a row from 0.0 in storage order.  `gather_rows`
reproduces exactly that order, so the routing results are bitwise equal to the
reference's (checked in oracle/validate_against_reference.py).
"""

import numpy as np

ROWOFF = (-1, -1, -1, 0, 0, 1, 1, 1)      # mrtm.py:144
COLOFF = (-1, 0, 1, -1, 1, -1, 0, 1)      # mrtm.py:145


def downstream(coords, flow_dir, nrow, ncol):
    """Downstream cell id (1-based; -1 = outlet) per cell, mrtm.py:85-120 + :233-258."""
    n = coords.shape[0]
    ids = coords[:, 0].astype(int)
    ilat = coords[:, 4].astype(int) - 1
    ilon = coords[:, 3].astype(int) - 1
    grid = np.zeros((nrow, ncol), dtype=int)
    grid[ilat, ilon] = ids

    fd = np.where(flow_dir == -9999., 0, flow_dir).astype(int)                 # :243-245
    rt, lt = 1 + 2 + 128, 8 + 16 + 32                                          # :237-240
    up, dn = 32 + 64 + 128, 2 + 4 + 8
    dlat = np.zeros(n, dtype=int)
    dlon = np.zeros(n, dtype=int)
    dlat[(dn & fd) != 0] = -1                                                  # :249-252 (later wins)
    dlat[(up & fd) != 0] = 1
    dlon[(rt & fd) != 0] = 1
    dlon[(lt & fd) != 0] = -1
    tlat = dlat + ilat
    tlon = dlon + ilon

    bad = (tlon < 0) | (tlon > ncol - 1)                                       # :100-102
    tlon[bad] = np.mod(tlon[bad] + 1, ncol)
    bad = (tlat < 0) | (tlat > nrow - 1)                                       # :104-106
    tlat[bad] = ilat[bad]
    tlon[bad] = ilon[bad]

    dsid = grid[tlat, tlon]
    dsid[(dsid == 0) | (dsid == ids)] = -1                                     # :116-118
    return dsid


def upstream(coords, dsid, nrow, ncol):
    """[N, 9]: 8 neighbour ids (inflowing first, stable) + inflow count, mrtm.py:123-191."""
    n = coords.shape[0]
    ids = coords[:, 0].astype(int)
    ilat = coords[:, 4].astype(int) - 1
    ilon = coords[:, 3].astype(int) - 1
    grid = np.zeros((nrow, ncol), dtype=int)
    grid[ilat, ilon] = ids

    nb = np.zeros((n, 8), dtype=int)
    isup = np.zeros((n, 8), dtype=bool)
    for k in range(8):
        r = ilat + ROWOFF[k]
        c = ilon + COLOFF[k]
        ok = (r >= 0) & (c >= 0) & (r <= nrow - 1) & (c <= ncol - 1)           # no wrap, :147-150
        nb[ok, k] = grid[r[ok], c[ok]]
        has = nb[:, k] != 0
        isup[has, k] = dsid[nb[has, k] - 1] == ids[has]                        # :164-166

    out = np.zeros((n, 9), dtype=int)
    for i in range(n):                                                         # stable partition, :169-184
        first = [nb[i, k] for k in range(8) if isup[i, k]]
        rest = [nb[i, k] for k in range(8) if not isup[i, k]]
        out[i, :8] = first + rest
        out[i, 8] = len(first)
    return out


def upstream_fast(coords, dsid, nrow, ncol):
    """Vectorised equivalent of `upstream` (same result, used for large grids)."""
    n = coords.shape[0]
    ids = coords[:, 0].astype(int)
    ilat = coords[:, 4].astype(int) - 1
    ilon = coords[:, 3].astype(int) - 1
    grid = np.zeros((nrow, ncol), dtype=int)
    grid[ilat, ilon] = ids
    nb = np.zeros((n, 8), dtype=int)
    isup = np.zeros((n, 8), dtype=bool)
    for k in range(8):
        r = ilat + ROWOFF[k]
        c = ilon + COLOFF[k]
        ok = (r >= 0) & (c >= 0) & (r <= nrow - 1) & (c <= ncol - 1)
        nb[ok, k] = grid[r[ok], c[ok]]
        has = nb[:, k] != 0
        isup[has, k] = dsid[nb[has, k] - 1] == ids[has]
    perm = np.argsort(~isup, axis=1, kind='stable')
    out = np.zeros((n, 9), dtype=int)
    out[:, :8] = np.take_along_axis(nb, perm, axis=1)
    out[:, 8] = isup.sum(axis=1)
    return out


def gather_rows(upid):
    """
    Rows of UM = UP - I (mrtm.py:194-230) as (cols [N, 9], sign [N, 9], count [N]):
    the upstream ids and the cell itself in ascending column order, padded.
    """
    n = upid.shape[0]
    cols = np.zeros((n, 9), dtype=np.int64)
    sign = np.zeros((n, 9), dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int64)
    for i in range(n):
        k = int(upid[i, 8])
        entries = [(int(j) - 1, 1.0) for j in upid[i, :k]] + [(i, -1.0)]
        entries.sort(key=lambda t: t[0])
        cnt[i] = k + 1
        for s, (j, w) in enumerate(entries):
            cols[i, s] = j
            sign[i, s] = w
    return cols, sign, cnt


def csr_rows(upid):
    """
    The same operator as a scipy CSR matrix with int64 data and sorted indices - what the reference's
    upstream_genmatrix returns after `.tocsr()` (mrtm.py:194-230).  scipy's csr_matvec accumulates a row from
    0.0 in ascending column order, i.e. exactly `_um_dot`; ~10x faster than the numpy gather at 67,420 cells,
    which is what makes full-size routing checks affordable.  Pass the result as `rows`.
    """
    import scipy.sparse as sparse
    cols, sign, cnt = gather_rows(upid)
    n = upid.shape[0]
    mask = np.arange(9)[None, :] < cnt[:, None]
    indptr = np.concatenate([[0], np.cumsum(cnt)])
    return sparse.csr_matrix((sign[mask].astype(np.int64), cols[mask], indptr), shape=(n, n))


def _um_dot(cols, sign, cnt, F):
    """Row-ordered accumulation from 0.0, like scipy's csr_matvec."""
    acc = np.zeros(F.shape[0])
    for s in range(9):
        live = cnt > s
        if not live.any():
            break
        acc[live] = acc[live] + sign[live, s] * F[cols[live, s]]
    return acc


def streamrouting(L, S0, F0, ChV, q, area, nday, dt, rows):
    """One month of routing, mrtm.py:16-82.  `rows` = gather_rows(upid)."""
    if isinstance(rows, tuple):
        cols, sign, cnt = rows
        um_dot = lambda f: _um_dot(cols, sign, cnt, f)                         # noqa: E731
    else:
        um_dot = rows.dot                                                      # scipy CSR, as the reference
    nt = int(nday * 24 * 3600 / dt)                                            # :36
    S = np.copy(S0)
    F = np.copy(F0)
    Favg = np.zeros(L.shape[0])
    tauinv = ChV / L                                                           # :42
    dtinv = 1. / dt
    erl = (q * area) * (1e6 / 1e3) / (nday * 24 * 3600)                        # :45
    for _ in range(nt):
        F = S * tauinv                                                         # :50
        dSdt = um_dot(F) + erl                                                 # :51
        Sx = (dSdt * dt) < (-S)                                                # :54
        if Sx.any():
            F[Sx] = dSdt[Sx] + F[Sx] + S[Sx] * dtinv                           # :60
            S[Sx] = 0                                                          # :63
            Sxn = np.logical_not(Sx)
            d2 = um_dot(F) + erl                                               # :68
            S[Sxn] += d2[Sxn] * dt
        else:
            S += (dSdt * dt)                                                   # :76
        Favg += F                                                              # :78
    Favg /= nt
    return S, Favg, F


def route(runoff, L, ChV, area, ndays, dt, rows, spinup_months, chs_prev=None):
    """
    Components.calculate_routing (components.py:262-296): spin-up over the first
    `spinup_months` months, then all months, continuing from the spun-up storage.
    runoff [N, M]; ndays [M] (mod-4 leap rule).  Returns ChStorage, Avg_ChFlow [N, M]
    and the final instantaneous flow [N].
    """
    n, m = runoff.shape
    S = np.zeros(n) if chs_prev is None else np.copy(chs_prev)
    F = np.zeros(n)
    chs = np.zeros((n, m))
    avg = np.zeros((n, m))
    for nm in range(spinup_months):
        S, _, F = streamrouting(L, S, F, ChV, runoff[:, nm], area, ndays[nm], dt, rows)
    for nm in range(m):
        S, Favg, F = streamrouting(L, S, F, ChV, runoff[:, nm], area, ndays[nm], dt, rows)
        chs[:, nm] = S
        avg[:, nm] = Favg
    return chs, avg, F
