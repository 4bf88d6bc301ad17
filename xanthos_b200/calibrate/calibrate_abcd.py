"""
Calibration of the ABCD model on the B200 - drop-in for xanthos/calibrate/calibrate_abcd.py.

Kept from the reference: `Calibrate` (same constructor, `calibrate_basin`, the two .npy result
files, :20-131), `basin_runoff` (:134-173), `objective_kge` (:176-213), `process_basin` (:216-232),
`expand_str_range` (:235-253) and `calibrate_all` (:256-262).

What changes is WHERE the objective is evaluated: one CUDA launch (`xan_abcd_kge_batch`) runs the
full spin-up + simulation + basin aggregation + KGE distance for every (basin, candidate) pair of a
differential-evolution generation, instead of one `ABCD.emulate()` per candidate.  The DE driver
follows scipy.optimize.differential_evolution's defaults as used by the reference (best1bin,
Latin-hypercube init, dither in [0.5, 1), recombination 0.7, tol 0.01, popsize 15 x n_params,
maxiter 1000, no polish) with *deferred* updating, so that a whole generation is one batch; all
basins of a call are advanced together.  Two drivers with the same semantics:
`differential_evolution_device` (default; population and generation logic on the device, csrc/de.cu)
and `differential_evolution_batched` (numpy, one host round trip per generation).

Targets: observed runoff (`set_calibrate = 0`, `BasinEvaluator`) and observed streamflow
(`set_calibrate = 1`, `StreamflowEvaluator`).  The reference's streamflow branch (:164-173) hands
the whole [ncell, nmonths] array returned by Components.calculate_routing to np.std / np.mean /
np.corrcoef (moments over every cell of the globe, correlation with global cell 0 after a
67,421 x 67,421 covariance matrix) and puts the basin's runoff at FLAT indices of the global
array (np.put): it has no usable behaviour to reproduce.  What is implemented is the INTENDED
objective of docs/calibration_tutorial.md - KGE between the routed flow at the basin's outlet
cell and the observed streamflow in m3/s - restated in oracle/calibrate.py
(`objective_kge_streamflow`, `outlet_cells`) and labelled as such.
"""

import logging
import time

import numpy as np

from .. import _cuda as C
from ..runoff import abcd as abcd_mod

LB = 1e-4                                                   # calibrate_abcd.py:62-67
BOUNDS_SNOW = [(LB, 1 - LB), (LB, 8 - LB), (LB, 1 - LB), (LB, 1 - LB), (LB, 1 - LB)]


class BasinEvaluator:
    """Device-resident forcing of a world + the batched KGE objective."""

    def __init__(self, basin_ids, basin_areas, precip, pet, tmin, n_months, runoff_spinup, obs_unit):
        torch = C.torch_cuda()
        self.basin_ids = np.asarray(basin_ids).astype(int)
        self.n_rows = int(self.basin_ids.max())
        rows = np.where(self.basin_ids >= 1, self.basin_ids - 1, -1).astype(np.int32)
        self.plan = abcd_mod.basin_plan(rows, self.n_rows)
        self.pet = C.as_field(pet)
        self.precip = C.as_field(precip)
        self.tmin = None if tmin is None else C.as_field(tmin)
        self.area = C.dev_vector(basin_areas)
        self.n_months = int(n_months)
        self.spinup = int(runoff_spinup)
        if obs_unit not in ('km3_per_mth', 'mm_per_mth'):
            raise C.ValidationException("obs_unit '{}' not supported for runoff calibration".format(obs_unit))
        self.unit_km3 = int(obs_unit == 'km3_per_mth')
        self.nosnow = tmin is None
        self._torch = torch

    def evaluate(self, basin_nums, pars, obs, want_series=False):
        """
        basin_nums [nb] (1-based ids), pars [nb, P, 4 or 5], obs [nb, n_months]
        -> KGE distance [nb, P] (and modelled basin series [nb, P, n_months]).
        """
        torch = self._torch
        pars = np.asarray(pars, dtype=np.float64)
        nb, npar, k = pars.shape
        if k == 4:
            pars = np.concatenate([pars, np.zeros((nb, npar, 1))], axis=2)
        d_pars = torch.from_numpy(np.ascontiguousarray(pars)).cuda()
        d_obs = torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float64)).cuda()
        d_series = torch.empty((nb, npar, self.n_months), dtype=torch.float64, device='cuda') if want_series else None
        ed = self.evaluate_device(basin_nums, d_pars, d_obs, d_series).cpu().numpy()
        if want_series:
            return ed, d_series.cpu().numpy()
        return ed

    def evaluate_device(self, basin_nums, d_pars, d_obs, d_series=None):
        """The same with parameters [nb, P, 5] and observations [nb, n_months] already on the device (cuda tensors);
        returns the KGE distances as a cuda tensor [nb, P].  Nothing is synchronised."""
        torch = self._torch
        nb, npar = int(d_pars.shape[0]), int(d_pars.shape[1])
        if tuple(d_obs.shape) != (nb, self.n_months) or len(basin_nums) != nb or int(d_pars.shape[2]) != 5:
            raise C.ValidationException("objective: parameters {} / observations {} do not match {} basins x {} months"
                                        .format(tuple(d_pars.shape), tuple(d_obs.shape), len(basin_nums), self.n_months))
        rows, rp = C.as_c(np.asarray(basin_nums, dtype=np.int64) - 1, np.int32)
        d_ed = torch.empty((nb, npar), dtype=torch.float64, device='cuda')
        C.check(C.lib().xan_abcd_kge_batch(self.plan._plan, rp, nb, npar, C.ptr(self.pet.t), C.ptr(self.precip.t),
                                           C.ptr(self.tmin.t if self.tmin is not None else None), C.ptr(self.area),
                                           C.ptr(d_pars), C.ptr(d_obs), self.n_months, self.spinup, self.pet.ld,
                                           self.unit_km3, C.ptr(d_ed), C.ptr(d_series), C.stream_ptr()))
        return d_ed


def outlet_cells(basin_ids, dsid, area):
    """
    Outlet cell (0-based) of every basin id 1 .. max: the basin's cell with the largest drainage area (own area plus
    everything upstream), lowest index on ties; -1 for ids without cells.  `dsid`: 1-based downstream id per cell,
    0 = none (routing/mrtm.py `downstream`).  The reference names no cell (module docstring): INTENDED semantics.
    """
    basin_ids = np.asarray(basin_ids).astype(np.int64)
    down = np.asarray(dsid).astype(np.int64) - 1
    n = len(down)
    acc = np.asarray(area, dtype=np.float64).copy()
    has = down >= 0
    indeg = np.bincount(down[has], minlength=n)
    frontier = np.nonzero(indeg == 0)[0]
    while len(frontier):               # leaves first, level by level; np.add.at adds the givers of a receiver in index order
        give = frontier[has[frontier]]
        recv = down[give]
        np.add.at(acc, recv, acc[give])
        np.subtract.at(indeg, recv, 1)
        cand = np.unique(recv)
        frontier = cand[indeg[cand] == 0]
    nb = int(basin_ids.max())
    out = np.full(nb, -1, dtype=np.int64)
    order = np.lexsort((np.arange(n), -acc))                # largest drainage area first, lowest index on ties
    ids = basin_ids[order]
    first = np.unique(ids, return_index=True)
    for b, i in zip(*first):
        if 1 <= b <= nb:
            out[b - 1] = order[i]
    return out


class StreamflowEvaluator:
    """
    KGE distance between the routed streamflow at every basin's outlet and its observed streamflow (m3/s) for a whole
    differential-evolution population: INTENDED semantics of `set_calibrate = 1` (module docstring).

    Routing never mixes basins, so ONE global pass evaluates candidate j of EVERY basin at once: for population slot j the
    parameter table row of basin b is candidate j of b, `xan_abcd_run` gives the global runoff field, and the fields of
    two slots share a launch of `xan_mrtm_route_batch` (members = population slots).  A generation of P candidates is
    P ABCD runs + P / 2 routing launches whatever the number of basins (0.5 degree world, 360 + 360 months, random
    candidates: 57.5 ms per population slot, tools/calib_streamflow_bench.py).
    Same `evaluate` / `evaluate_device` interface as `BasinEvaluator`.
    """

    def __init__(self, basin_ids, basin_areas, precip, pet, tmin, n_months, runoff_spinup, um, dsid, flow_dist,
                 str_velocity, ndays, dt, routing_spinup, obs_unit='m3_per_sec'):
        from ..routing import mrtm as mrtm_mod
        torch = C.torch_cuda()
        if obs_unit != 'm3_per_sec':
            raise C.ValidationException("obs_unit '{}' not supported for streamflow calibration".format(obs_unit))
        self.basin_ids = np.asarray(basin_ids).astype(int)
        self.n_rows = int(self.basin_ids.max())
        rows = np.where(self.basin_ids >= 1, self.basin_ids - 1, -1).astype(np.int32)
        self.plan = abcd_mod.basin_plan(rows, self.n_rows)
        self.pet, self.precip = C.as_field(pet), C.as_field(precip)
        self.tmin = None if tmin is None else C.as_field(tmin)
        self.nosnow = tmin is None
        self.n_months, self.spinup = int(n_months), int(runoff_spinup)
        self.um, self._mrtm = um, mrtm_mod
        self.d_L, self.d_V, self.d_A = C.dev_vector(flow_dist), C.dev_vector(str_velocity), C.dev_vector(basin_areas)
        self.ndays = np.asarray(ndays, dtype=np.int32).reshape(-1)[:self.n_months]
        self.dt, self.routing_spinup = float(dt), int(routing_spinup)
        self.outlets = outlet_cells(self.basin_ids, dsid, basin_areas)
        self.d_outlets = torch.from_numpy(np.maximum(self.outlets, 0)).cuda()
        self._torch = torch

    def evaluate(self, basin_nums, pars, obs, want_series=False):
        """basin_nums [nb] (1-based), pars [nb, P, 4 or 5], obs [nb, n_months] -> distances [nb, P] (+ series)."""
        torch = self._torch
        pars = np.asarray(pars, dtype=np.float64)
        nb, npar, k = pars.shape
        if k == 4:
            pars = np.concatenate([pars, np.zeros((nb, npar, 1))], axis=2)
        d_pars = torch.from_numpy(np.ascontiguousarray(pars)).cuda()
        d_obs = torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float64)).cuda()
        d_series = torch.empty((nb, npar, self.n_months), dtype=torch.float64, device='cuda') if want_series else None
        ed = self.evaluate_device(basin_nums, d_pars, d_obs, d_series).cpu().numpy()
        return (ed, d_series.cpu().numpy()) if want_series else ed

    def evaluate_device(self, basin_nums, d_pars, d_obs, d_series=None):
        torch = self._torch
        nb, npar = int(d_pars.shape[0]), int(d_pars.shape[1])
        if tuple(d_obs.shape) != (nb, self.n_months) or len(basin_nums) != nb or int(d_pars.shape[2]) != 5:
            raise C.ValidationException("objective: parameters {} / observations {} do not match {} basins x {} months"
                                        .format(tuple(d_pars.shape), tuple(d_obs.shape), len(basin_nums), self.n_months))
        rows_h = np.asarray(basin_nums, dtype=np.int64) - 1
        if (rows_h < 0).any() or (rows_h >= self.n_rows).any() or (self.outlets[rows_h] < 0).any():
            raise C.ValidationException("streamflow calibration: a requested basin has no cells")
        rows = torch.from_numpy(rows_h).cuda()
        # rows of the basins that are not being calibrated: any valid parameter set (their flow is not looked at)
        base = torch.tensor([0.5, 4.0, 0.5, 0.5, 0.5], dtype=torch.float64, device='cuda').repeat(self.n_rows, 1)
        outl = self.d_outlets[rows]
        mod = d_series if d_series is not None else torch.empty((nb, npar, self.n_months), dtype=torch.float64, device='cuda')
        for j0 in range(0, npar, 2):
            js = list(range(j0, min(npar, j0 + 2)))
            qs = []
            for j in js:
                table = base.clone()
                table[rows] = d_pars[:, j, :]
                if self.nosnow:
                    table = table[:, :4].contiguous()
                qs.append(abcd_mod.run_device(self.plan, table, self.pet, self.precip, self.tmin, self.n_months,
                                              self.spinup, want=('q',))['q'])
            if len(qs) == 1:
                routed = [self._mrtm.route_device(self.um, qs[0], self.d_L, self.d_V, self.d_A, self.ndays, self.dt,
                                                  self.routing_spinup, want_chs=False)]
            else:
                routed = self._mrtm.route_device_batch(self.um, qs, self.d_L, self.d_V, self.d_A, self.ndays, self.dt,
                                                       self.routing_spinup, want_chs=False)
            for j, (_, avg, _) in zip(js, routed):
                mod[:, j, :] = avg.t[:, outl].t()             # Avg_ChFlow [nmonths][ld] -> outlet series [nb, n_months]
        return kge_distance_device(mod, d_obs)


def kge_distance_device(mod, obs):
    """Euclidean distance from the KGE optimum (calibrate_abcd.py:190-211) for mod [nb, P, M] against obs [nb, M]:
    population standard deviations and means (np.std / np.mean), Pearson correlation (np.corrcoef)."""
    torch = C.torch_cuda()
    o = obs[:, None, :]
    m_m, m_o = mod.mean(dim=2), o.mean(dim=2)
    dm, do = mod - m_m[..., None], o - m_o[..., None]
    var_m, var_o = (dm * dm).mean(dim=2), (do * do).mean(dim=2)
    relvar = torch.sqrt(var_m) / torch.sqrt(var_o)
    bias = m_m / m_o
    r = (dm * do).mean(dim=2) / torch.sqrt(var_m * var_o)
    return torch.sqrt((r - 1) ** 2 + (relvar - 1) ** 2 + (bias - 1) ** 2)


def differential_evolution_batched(evaluate, n_problems, bounds, popsize=15, maxiter=1000, tol=0.01, atol=0.0,
                                   mutation=(0.5, 1.0), recombination=0.7, seed=None, callback=None):
    """
    best1bin differential evolution for `n_problems` independent problems advanced in lock-step.

    evaluate(x [k, S, D], idx [k]) -> energies [k, S] for the problems `idx` (lower is better; NaN
    counts as +inf).
    Follows scipy's solver (scaled [0, 1] parameters, Latin-hypercube init, per-generation dither,
    binomial crossover with one forced mutant gene, out-of-bounds genes re-drawn uniformly,
    convergence when std(E) <= atol + tol * |mean(E)|) with deferred updating.
    Returns dict(x [n, D], fun [n], nit [n], nfev [n]).
    """
    rng = np.random.default_rng(seed)
    bounds = np.asarray(bounds, dtype=float)
    D = bounds.shape[0]
    S = max(5, popsize * D)
    lo, span = bounds[:, 0], bounds[:, 1] - bounds[:, 0]
    n = n_problems

    # Latin hypercube in [0, 1]^D per problem
    seg = 1.0 / S
    pop = np.empty((n, S, D))
    for j in range(D):
        samples = seg * rng.random((n, S)) + np.linspace(0.0, 1.0, S, endpoint=False)[None, :]
        perm = np.argsort(rng.random((n, S)), axis=1)
        pop[:, :, j] = np.take_along_axis(samples, perm, axis=1)

    def energies(p, which):
        e = np.asarray(evaluate(lo + p * span, which), dtype=float)
        return np.where(np.isnan(e), np.inf, e)

    E = energies(pop, np.arange(n))
    nfev = np.full(n, S)
    nit = np.zeros(n, dtype=int)
    active = np.ones(n, dtype=bool)

    def converged(e):
        with np.errstate(invalid='ignore'):
            return np.std(e, axis=1) <= atol + tol * np.abs(np.mean(e, axis=1))

    active &= ~(np.all(np.isfinite(E), axis=1) & converged(E))
    idx = np.arange(S)
    for it in range(1, maxiter + 1):
        if not active.any():
            break
        act = np.nonzero(active)[0]
        na = len(act)
        P, Ea = pop[act], E[act]
        best = P[np.arange(na), np.argmin(Ea, axis=1)]                      # [na, D]
        scale = rng.uniform(mutation[0], mutation[1], size=(na, 1, 1))       # dither, one draw per generation
        # two distinct random members different from i
        r0 = (idx[None, :] + rng.integers(1, S, size=(na, S))) % S
        r1 = (idx[None, :] + rng.integers(1, S - 1, size=(na, S))) % S
        clash = r1 == r0
        r1 = np.where(clash, (r1 + 1) % S, r1)
        r1 = np.where(r1 == idx[None, :], (r1 + 1) % S, r1)
        r1 = np.where(r1 == r0, (r1 + 1) % S, r1)
        a = np.take_along_axis(P, r0[:, :, None], axis=1)
        b = np.take_along_axis(P, r1[:, :, None], axis=1)
        mutant = best[:, None, :] + scale * (a - b)
        cross = rng.random((na, S, D)) < recombination
        forced = rng.integers(0, D, size=(na, S))
        cross[np.arange(na)[:, None], idx[None, :], forced] = True
        trial = np.where(cross, mutant, P)
        oob = (trial < 0) | (trial > 1)
        trial = np.where(oob, rng.random((na, S, D)), trial)
        Et = energies(trial, act)
        better = Et <= Ea
        P = np.where(better[:, :, None], trial, P)
        Ea = np.where(better, Et, Ea)
        pop[act], E[act] = P, Ea
        nfev[act] += S
        nit[act] = it
        done = np.all(np.isfinite(Ea), axis=1) & converged(Ea)
        active[act[done]] = False
        if callback is not None:
            callback(it, E)
    b = np.argmin(E, axis=1)
    return dict(x=lo + pop[np.arange(n), b] * span, fun=E[np.arange(n), b], nit=nit, nfev=nfev)


def differential_evolution_device(ev, basin_nums, robs, bounds, popsize=15, maxiter=1000, tol=0.01, atol=0.0,
                                  mutation=(0.5, 1.0), recombination=0.7, seed=None, check_every=4, pop_members=None):
    """
    The same solver with population, energies and generation logic on the device (xan_de_init / xan_de_trial /
    xan_de_select, csrc/de.cu): a generation is trial kernel -> objective kernels -> selection kernel, and the host
    only reads the convergence flags every `check_every` generations to drop finished basins from the batch (a
    converged basin is frozen on the device at once, so the result does not depend on `check_every`).
    ev: BasinEvaluator; robs [n, n_months].  Returns dict(x [n, D], fun [n], nit [n], nfev [n]).
    """
    torch = C.torch_cuda()
    lib = C.lib()
    bounds = np.asarray(bounds, dtype=float)
    D = bounds.shape[0]
    S = max(5, popsize * D) if pop_members is None else int(pop_members)   # scipy: popsize x D members
    n = len(basin_nums)
    bn = np.asarray(basin_nums)
    seed = int(np.random.SeedSequence(seed).generate_state(2, dtype=np.uint32).astype(np.uint64) @ np.array([1, 1 << 32], dtype=np.uint64))
    dev = dict(dtype=torch.float64, device='cuda')
    d_lo = torch.tensor(bounds[:, 0], **dev)
    d_span = torch.tensor(bounds[:, 1] - bounds[:, 0], **dev)
    d_obs = torch.from_numpy(np.ascontiguousarray(robs, dtype=np.float64)).cuda()
    pop = torch.empty((n, S, D), **dev)
    conv = torch.zeros(n, dtype=torch.int32, device='cuda')
    C.check(lib.xan_de_init(C.ptr(pop), n, S, D, seed, C.stream_ptr()))
    pars0 = torch.zeros((n, S, 5), **dev)
    pars0[:, :, :D] = d_lo + pop * d_span
    E = ev.evaluate_device(bn, pars0, d_obs)
    act_h = np.arange(n, dtype=np.int32)
    act = torch.from_numpy(act_h).cuda()
    C.check(lib.xan_de_select(C.ptr(pop), C.ptr(E), C.ptr(act), n, S, D, None, None, float(tol), float(atol), 0,
                              C.ptr(conv), C.stream_ptr()))
    nfev = np.full(n, S)
    nit = np.zeros(n, dtype=int)

    def refresh_active():
        c = conv.cpu().numpy()
        return np.nonzero(c == 0)[0].astype(np.int32), c
    act_h, c_h = refresh_active()
    obs_a = d_obs
    it = 0
    while it < maxiter and len(act_h):
        na = len(act_h)
        act = torch.from_numpy(act_h).cuda()
        obs_a = d_obs.index_select(0, act.long()) if na < n else d_obs
        trial_x = torch.empty((na, S, D), **dev)
        trial_p = torch.empty((na, S, 5), **dev)
        for _ in range(min(check_every, maxiter - it)):
            it += 1
            C.check(lib.xan_de_trial(C.ptr(pop), C.ptr(E), C.ptr(act), na, S, D, 5, C.ptr(d_lo), C.ptr(d_span), seed, it,
                                     float(mutation[0]), float(mutation[1]), float(recombination), C.ptr(trial_x),
                                     C.ptr(trial_p), C.stream_ptr()))
            Et = ev.evaluate_device(bn[act_h], trial_p, obs_a)
            C.check(lib.xan_de_select(C.ptr(pop), C.ptr(E), C.ptr(act), na, S, D, C.ptr(trial_x), C.ptr(Et), float(tol),
                                      float(atol), it, C.ptr(conv), C.stream_ptr()))
        prev = act_h
        act_h, c_h = refresh_active()
        done_gen = np.where(c_h[prev] > 0, c_h[prev], it)          # generation at which each problem stopped
        nit[prev] = done_gen
        nfev[prev] = S * (1 + done_gen)
    E_h = E.cpu().numpy()
    pop_h = pop.cpu().numpy()
    b = np.argmin(np.where(np.isnan(E_h), np.inf, E_h), axis=1)
    return dict(x=bounds[:, 0] + pop_h[np.arange(n), b] * (bounds[:, 1] - bounds[:, 0]), fun=E_h[np.arange(n), b],
                nit=nit, nfev=nfev)


def streamflow_evaluator(router_function, basin_ids, basin_areas, precip, pet, tmin, n_months, runoff_spinup,
                         obs_unit='m3_per_sec'):
    """
    StreamflowEvaluator for the world behind `router_function` - the bound `Components.calculate_routing` that the
    reference passes down as router_func (components.py:497, calibrate_abcd.py:216-232): its instance holds the flow
    network, the channel geometry, the calendar and the routing time step.
    """
    from ..routing import mrtm as mrtm_mod
    comp = getattr(router_function, '__self__', None)
    if comp is None or not hasattr(comp, 'data') or not hasattr(comp, 's'):
        raise C.ValidationException("streamflow calibration (set_calibrate = 1) needs router_function = "
                                    "Components.calculate_routing of the run being calibrated")
    if getattr(comp, 'um', None) is None or getattr(comp, 'dsid', None) is None:
        comp.dsid = mrtm_mod.downstream(comp.data.coords, comp.data.flow_dir, comp.s)
        comp.upid = mrtm_mod.upstream(comp.data.coords, comp.dsid, comp.s)
        comp.um = mrtm_mod.upstream_genmatrix(comp.upid)
    return StreamflowEvaluator(basin_ids, basin_areas, precip, pet, tmin, n_months, runoff_spinup, comp.um, comp.dsid,
                               comp.data.flow_dist, comp.data.str_velocity, comp.yr_imth_dys[:, 2],
                               comp.routing_timestep_hours, comp.s.routing_spinup, obs_unit)


def _basin_obs(obs, basin_num, n_months):
    """Observed series of a basin (calibrate_abcd.py:88)."""
    return np.asarray(obs)[np.where(np.asarray(obs)[:, 0] == basin_num)][:n_months, 1]


def calibrate_basins(basin_nums, basin_ids, basin_areas, precip, pet, obs, tmin, n_months, runoff_spinup, obs_unit,
                     popsize=15, maxiter=1000, tol=0.01, seed=None, evaluator=None, driver='device'):
    """
    Differential evolution for several basins at once.  Returns (pars [nb, 4 or 5], kge [nb], info).
    """
    ev = evaluator or BasinEvaluator(basin_ids, basin_areas, precip, pet, tmin, n_months, runoff_spinup, obs_unit)
    basin_nums = [int(b) for b in basin_nums]
    series = [_basin_obs(obs, b, n_months) for b in basin_nums]
    short = [b for b, r in zip(basin_nums, series) if r.shape[0] != int(n_months)]
    if short:       # the reference fails in np.corrcoef on the length mismatch (calibrate_abcd.py:200)
        raise C.ValidationException("observations of basin(s) {} do not cover the {} simulated months".format(
            short, int(n_months)))
    robs = np.stack(series)
    bounds = BOUNDS_SNOW[:4] if ev.nosnow else BOUNDS_SNOW
    bn = np.asarray(basin_nums)
    if driver == 'host':       # numpy generation logic, one host round trip per generation
        res = differential_evolution_batched(lambda x, idx: ev.evaluate(bn[idx], x, robs[idx]), len(basin_nums), bounds,
                                             popsize=popsize, maxiter=maxiter, tol=tol, seed=seed)
    else:
        res = differential_evolution_device(ev, bn, robs, bounds, popsize=popsize, maxiter=maxiter, tol=tol, seed=seed)
    return res['x'], 1 - res['fun'], res


class Calibrate:
    """Calibrate the ABCD runoff module (same constructor as the reference, calibrate_abcd.py:23-88)."""

    def __init__(self, basin_num, basin_ids, basin_areas, precip, pet, obs, tmin, n_months, runoff_spinup,
                 set_calibrate, obs_unit, out_dir, router_func=None):
        if set_calibrate not in (0, 1):
            raise C.ValidationException("set_calibrate must be 0 (observed runoff) or 1 (observed streamflow)")
        self.basin_num = basin_num
        self.basin_ids = basin_ids
        self.basin_areas = basin_areas
        self.precip = precip
        self.pet = pet
        self.obs = obs
        self.tmin = tmin
        self.n_months = n_months
        self.runoff_spinup = runoff_spinup
        self.router_func = router_func
        self.set_calibrate = set_calibrate
        self.obs_unit = obs_unit
        self.out_dir = out_dir
        self.nosnow = self.tmin is None
        self.bounds = list(BOUNDS_SNOW[:4] if self.nosnow else BOUNDS_SNOW)
        self.all_pars = np.zeros((1, len(self.bounds)))
        self.kge_vals = np.zeros(1)
        self.basin_idx = np.where(self.basin_ids == self.basin_num)
        self.bsn_areas = self.basin_areas[self.basin_idx]
        self.bsn_Robs = _basin_obs(self.obs, basin_num, self.n_months)

    def calibrate_basin(self, popsize=15, polish=False, seed=None, maxiter=1000):
        """Calibrate the basin and save kge_result_basin_<n>.npy / abcd(m)_parameters_basin_<n>.npy (:90-131)."""
        st = time.time()
        ev = None
        if self.set_calibrate == 1:      # routed flow at the basin's outlet against observed streamflow (module docstring)
            ev = streamflow_evaluator(self.router_func, self.basin_ids, self.basin_areas, self.precip, self.pet, self.tmin,
                                      self.n_months, self.runoff_spinup, self.obs_unit)
        pars, kge, res = calibrate_basins([self.basin_num], self.basin_ids, self.basin_areas, self.precip, self.pet,
                                          self.obs, self.tmin, self.n_months, self.runoff_spinup, self.obs_unit,
                                          popsize=popsize, maxiter=maxiter, seed=seed, evaluator=ev)
        self.all_pars[0, :] = pars[0]
        self.kge_vals[0] = kge[0]
        par_names = 'abcd' + 'm' * (not self.nosnow)
        logging.debug("\t\tFinished calibration for basin {0} which contains {1} grid cells.".format(
            self.basin_num, self.basin_idx[0].shape[0]))
        logging.debug("\t\tPopulation size:  {}".format(popsize))
        logging.debug("\t\tParameter values ({}):  {}".format(','.join(list(par_names)), pars[0]))
        logging.debug("\t\tKGE:  {}".format(kge[0]))
        logging.debug("\t\tNumber of function evaluations:  {}".format(int(res['nfev'][0])))
        logging.debug("\t\tCalibration time (seconds):  {}".format(time.time() - st))
        save_results(self.out_dir, self.basin_num, self.kge_vals, self.all_pars, par_names)


def save_results(out_dir, basin_num, kge_vals, all_pars, par_names):
    np.save('{}/kge_result_basin_{}.npy'.format(out_dir, basin_num), kge_vals)
    np.save('{}/{}_parameters_basin_{}.npy'.format(out_dir, par_names, basin_num), all_pars)


def _one_basin_world(set_calibrate, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas, basin_idx, arr_shp,
                     routing_func):
    """Evaluator and basin number for the reference-signature calls below (one basin, one parameter vector)."""
    idx = np.asarray(basin_idx[0] if isinstance(basin_idx, tuple) else basin_idx)
    n = len(idx)
    if set_calibrate == 0:
        return BasinEvaluator(np.ones(n, dtype=int), bsn_areas, precip, pet, tmin, n_months, runoff_spinup, obs_unit), 1
    # streamflow target: the basin's rows go back into global arrays (as the reference does with the runoff, :169-171);
    # the other cells get zero forcing - their flow stays in their own basins and is not looked at
    comp = getattr(routing_func, '__self__', None)
    if comp is None:
        raise C.ValidationException("streamflow calibration needs routing_func = Components.calculate_routing")
    ncell = int(arr_shp[0])

    def spread(a):
        g = np.zeros((ncell, np.asarray(a).shape[1]))
        g[idx] = a
        return g
    ids = np.asarray(comp.data.basin_ids).astype(int)
    ev = streamflow_evaluator(routing_func, ids, comp.data.area, spread(precip), spread(pet),
                              None if tmin is None else spread(tmin), n_months, runoff_spinup, obs_unit)
    return ev, int(ids[idx[0]])


def basin_runoff(pars, set_calibrate, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas, basin_idx,
                 arr_shp, routing_func=None):
    """Modelled series of one basin for one parameter vector (calibrate_abcd.py:134-173): the basin's runoff for the
    runoff target, the routed flow at its outlet (INTENDED semantics, module docstring) for the streamflow target."""
    ev, b = _one_basin_world(set_calibrate, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas, basin_idx,
                             arr_shp, routing_func)
    _, series = ev.evaluate([b], np.asarray(pars, dtype=float)[None, None, :], np.ones((1, int(n_months))),
                            want_series=True)
    return series[0, 0]


def objective_kge(pars, model_func, set_calibrate, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas,
                  bsn_Robs, basin_idx, arr_shp, routing_func=None):
    """KGE distance between simulated and observed basin series (calibrate_abcd.py:176-213)."""
    ev, b = _one_basin_world(set_calibrate, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas, basin_idx,
                             arr_shp, routing_func)
    ed = ev.evaluate([b], np.asarray(pars, dtype=float)[None, None, :], np.asarray(bsn_Robs, dtype=float)[None, :])
    return float(ed[0, 0])


def process_basin(basin_num, settings, data, pet, router_function=None):
    """Process single basin (calibrate_abcd.py:216-232)."""
    cal = Calibrate(basin_num=basin_num, set_calibrate=settings.set_calibrate, obs_unit=settings.obs_unit,
                    basin_ids=data.basin_ids, basin_areas=data.area, precip=data.precip, pet=pet, obs=data.cal_obs,
                    tmin=data.tmin, n_months=settings.nmonths, runoff_spinup=settings.runoff_spinup,
                    router_func=router_function, out_dir=settings.calib_out_dir)
    cal.calibrate_basin()


def expand_str_range(str_ranges):
    """['0-2', '6', '7-9'] -> [0, 1, 2, 6, 7, 8, 9] (calibrate_abcd.py:235-253)."""
    out_list = []
    for r in str_ranges:
        if '-' in r:
            start, end = r.split('-')
            out_list.extend(range(int(start), int(end) + 1))
        else:
            out_list.append(int(r))
    return out_list


def calibrate_all(settings, data, pet, router_function, popsize=15, maxiter=1000, seed=None):
    """
    Calibrate all requested basins (calibrate_abcd.py:256-262).  The reference loops over the basins
    one differential evolution at a time; here all of them advance together, one CUDA launch per
    generation, and the same per-basin result files are written.
    """
    basins = expand_str_range(settings.cal_basins)
    for b in basins:
        name = data.basin_names[b - 1] if getattr(data, 'basin_names', None) is not None else ''
        logging.info("\tCalibrating Basin:  {} ({})".format(b, name))
    # One process per GPU (torch.distributed initialised by the launcher): whole basins per rank, balanced by cell
    # count - every basin's differential evolution is independent (the reference's loop, :256-262, has no coupling
    # either).  Each rank writes the result files of its own basins; the parameter / KGE tables are gathered.
    import torch.distributed as dist
    dist_on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist_on else (0, 1)
    mine = basins
    if dist_on:
        from .. import sharding
        mine = [int(b) for b in sharding.partition_basins(data.basin_ids, world, basins=basins)[rank]]
    st = time.time()
    par_names = 'abcd' + 'm' * (data.tmin is not None)
    npar = len(par_names)
    pars, kge, nfev = np.zeros((0, npar)), np.zeros(0), 0
    if mine:
        ev = None
        if settings.set_calibrate == 1:   # routed flow at every basin's outlet against observed streamflow (module docstring)
            ev = streamflow_evaluator(router_function, data.basin_ids, data.area, data.precip, pet, data.tmin,
                                      settings.nmonths, settings.runoff_spinup, settings.obs_unit)
        pars, kge, res = calibrate_basins(mine, data.basin_ids, data.area, data.precip, pet, data.cal_obs, data.tmin,
                                          settings.nmonths, settings.runoff_spinup, settings.obs_unit, popsize=popsize,
                                          maxiter=maxiter, seed=seed, evaluator=ev)
        nfev = int(res['nfev'].sum())
        for i, b in enumerate(mine):
            save_results(settings.calib_out_dir, b, np.array([kge[i]]), pars[i][None, :], par_names)
    if dist_on:
        import torch
        from .. import sharding
        pos = {b: i for i, b in enumerate(basins)}
        rows = [pos[b] for b in mine]
        local = torch.from_numpy(np.concatenate([pars, kge[:, None]], axis=1))
        local = local.cuda() if dist.get_backend() == 'nccl' else local
        full = sharding.gather_ragged_rows(rows, local, len(basins)).cpu().numpy()
        pars, kge = full[:, :npar], full[:, npar]
    logging.info("\tCalibration of {} basins ({} on this rank): {} evaluations in {:.2f} s".format(
        len(basins), len(mine), nfev, time.time() - st))
    return pars, kge
