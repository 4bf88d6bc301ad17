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

Only the runoff target (`set_calibrate = 0`) is supported: the streamflow branch of the
reference (:164-173) is broken.
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
        if set_calibrate != 0:
            raise NotImplementedError("only calibration against observed runoff (set_calibrate = 0) is supported; "
                                      "the reference's streamflow branch is broken (calibrate_abcd.py:164-173)")
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
        pars, kge, res = calibrate_basins([self.basin_num], self.basin_ids, self.basin_areas, self.precip, self.pet,
                                          self.obs, self.tmin, self.n_months, self.runoff_spinup, self.obs_unit,
                                          popsize=popsize, maxiter=maxiter, seed=seed)
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


def basin_runoff(pars, set_calibrate, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas, basin_idx,
                 arr_shp, routing_func=None):
    """Modelled runoff series of one basin for one parameter vector (calibrate_abcd.py:134-162)."""
    if set_calibrate != 0:
        raise NotImplementedError("streamflow calibration is not supported (broken in the reference)")
    n = len(basin_idx[0]) if isinstance(basin_idx, tuple) else len(basin_idx)
    ev = BasinEvaluator(np.ones(n, dtype=int), bsn_areas, precip, pet, tmin, n_months, runoff_spinup, obs_unit)
    _, series = ev.evaluate([1], np.asarray(pars, dtype=float)[None, None, :], np.ones((1, int(n_months))),
                            want_series=True)
    return series[0, 0]


def objective_kge(pars, model_func, set_calibrate, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas,
                  bsn_Robs, basin_idx, arr_shp, routing_func=None):
    """KGE distance between simulated and observed basin runoff (calibrate_abcd.py:176-213)."""
    if set_calibrate != 0:
        raise NotImplementedError("streamflow calibration is not supported (broken in the reference)")
    n = len(basin_idx[0]) if isinstance(basin_idx, tuple) else len(basin_idx)
    ev = BasinEvaluator(np.ones(n, dtype=int), bsn_areas, precip, pet, tmin, n_months, runoff_spinup, obs_unit)
    ed = ev.evaluate([1], np.asarray(pars, dtype=float)[None, None, :], np.asarray(bsn_Robs, dtype=float)[None, :])
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
    if settings.set_calibrate != 0:
        raise NotImplementedError("only calibration against observed runoff (set_calibrate = 0) is supported")
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
        pars, kge, res = calibrate_basins(mine, data.basin_ids, data.area, data.precip, pet, data.cal_obs, data.tmin,
                                          settings.nmonths, settings.runoff_spinup, settings.obs_unit, popsize=popsize,
                                          maxiter=maxiter, seed=seed)
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
