#!/usr/bin/env python
"""
Benchmark of the Xanthos hot path (BASELINE.json: cell-months/s for PM + ABCD + MRTM).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload = "pm_abcd_mrtm_360"): BASELINE.json configs[0]/[1] shape - Penman-Monteith
PET -> ABCD runoff -> MRTM routing, 67,420 cells x 360 months (1971-2000), synthetic forcing,
runoff spin-up 360 months, routing spin-up 360 months, 3-hour routing sub-steps, nlcs = 8.
A "step" is one pass of that pipeline over MEMBERS_PER_STEP = 2 scenario members of that shape (PM -> ABCD per
member, ONE routing launch for both: their thread blocks share the SMs); `single_member_step` reports the
one-member step of round 1.  At N > 1 every rank runs its own members (ensemble sharding, weak scaling, no
collective in the data path; the basin-aggregated runoff and streamflow [360 x 235] of every member are
all-gathered over NCCL at the end of the step).

  value : cell-months/s (members x cells x months / step time) with the forcing already resident in HBM
          (month-major fields), device time
  e2e   : same metric through the public ensemble API (xanthos_b200.ensemble.run_ensemble): every member's forcing
          from pinned HOST buffers, its outputs (q, avgchflow, basin aggregates) back to host ndarrays, H2D and D2H
          inside the timing, median of five runs; e2e.single_member = the reference-facing plug-in calls
          (run_pmpet / abcd_execute / route) on one member, nothing overlapped
  roofline, cpu_baseline, clocks, gpu_launches : as the measurement contract asks
  calib    : param-sets/s of one differential-evolution generation (235 basins x 64 candidates) and of a whole
             DE loop with the device and the numpy driver
  postproc : drought thresholds / statistics and basin sums on the step's resident runoff, numpy port beside it
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCELL, NMONTHS, START_YR, END_YR = 67420, 360, 1971, 2000
NLCS, N_BASINS = 8, 235
RUNOFF_SPINUP, ROUTING_SPINUP, DT = 360, 360, 3 * 3600
CALIB_POP = 64
WORKLOAD = "pm_abcd_mrtm_360"

# algorithmic bytes per cell-month (fp64, every array touched once; SURVEY.md section 8d / DESIGN.md)
BYTES_PM = 6 * 8 + NLCS * 8 / 12.0 + 8          # 6 forcings + land cover (once per year) + PET out = 61.3
BYTES_ABCD = 24 + 24 + 24 * RUNOFF_SPINUP / NMONTHS   # sim reads + writes + spin-up reads = 72
BYTES_MRTM = 8 + 16 + 8 * ROUTING_SPINUP / NMONTHS    # q + (ChStorage, Avg_ChFlow) + spin-up reads = 32


MEMBERS_PER_STEP = 2      # scenario members per device-resident step (one routing launch for both)


def bench_config(ncell, nmonths):
    """`config` of the JSON line - the SAME dictionary for both arms (`--impl ours` / `--impl reference`)."""
    spin_ro, spin_rt = min(RUNOFF_SPINUP, nmonths), min(ROUTING_SPINUP, nmonths)
    return {'workload': WORKLOAD if (ncell, nmonths) == (NCELL, NMONTHS) else 'reduced_%dx%d' % (ncell, nmonths),
            'ncell': ncell, 'nmonths': nmonths, 'nlcs': NLCS, 'runoff_spinup': spin_ro, 'routing_spinup': spin_rt,
            'dt_s': DT, 'members_per_gpu': MEMBERS_PER_STEP, 'members_per_step': MEMBERS_PER_STEP,
            'parallelism': 'members-per-gpu',
            'batch': 'a step is one pass of PM -> ABCD -> MRTM over %d scenario members of the named configuration (the '
                     'thread blocks of the members share the SMs in the routing launch); value = members x cells x '
                     'months / step time; the one-member step is reported as `single_member_step`' % MEMBERS_PER_STEP,
            'l2': 'inputs (8 x %.0f MB per member) exceed the 126 MB L2; no flush needed' % (ncell * nmonths * 8 / 1e6)}


def _peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.path, self.skip = index, None, None, 0

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """Start of the timed region: samples taken before it are dropped.  nvidia-smi is started BEFORE the warm-up
        because its start-up (NVML initialisation over all GPUs) stalls kernel launches for tens of milliseconds - inside
        the timed region that showed up as 60 ms instead of 52 ms per step in one run out of four."""
        if self.proc is None:
            return
        t0 = time.time()
        while time.time() - t0 < 5.0:            # wait for the first sample: nvidia-smi is up and running
            try:
                with open(self.path) as f:
                    self.skip = sum(1 for _ in f)
            except Exception:
                self.skip = 0
            if self.skip > 0:
                break
            time.sleep(0.05)

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for k, line in enumerate(open(self.path)):
                if k < self.skip:
                    continue
                p = [x.strip() for x in line.split(',')]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out['sm_mhz'] = float(np.median(sm))
            out['sm_max_mhz'] = float(max(smax))
        out['reasons'] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def build_inputs(member_seed=1, ncell=NCELL, nmonths=NMONTHS):
    from xanthos_b200 import synthetic
    if ncell == NCELL:
        world = synthetic.make_world(seed=0)
    else:
        world = synthetic.make_world(36, 72, ncell, 12, seed=0)
    end_yr = START_YR + nmonths // 12 - 1
    pm = synthetic.pm_inputs(world, START_YR, end_yr, nlcs=NLCS, seed=member_seed)
    ab = synthetic.abcd_inputs(world, nmonths, seed=member_seed, with_pet=False)
    for k in ('tair_load', 'TMIN_load', 'rhs_load', 'wind_load', 'rsds_load', 'rlds_load'):
        pm[k] = np.nan_to_num(pm[k])
    ab['tmin'] = np.nan_to_num(ab['tmin'])
    # The forcing VALUES are single precision, as in the NetCDF files of the climate models (float variables); the
    # arrays are float64 like the ones the reference's loader hands out.  Both arms compute on exactly these values.
    for k in ('tair_load', 'TMIN_load', 'rhs_load', 'wind_load', 'rsds_load', 'rlds_load'):
        pm[k] = pm[k].astype(np.float32).astype(np.float64)
    for k in ('precip', 'tmin'):
        ab[k] = ab[k].astype(np.float32).astype(np.float64)
    return world, pm, ab, end_yr


def month_days_mod4(nmonths, start_yr):
    d = []
    for y in range(start_yr, start_yr + nmonths // 12):
        d += [31, 29 if y % 4 == 0 else 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    return np.array(d, dtype=np.int32)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world_size, local_rank):
    import torch
    import torch.distributed as dist
    from xanthos_b200 import _cuda as C
    from xanthos_b200.pet import penman_monteith as pm_mod
    from xanthos_b200.runoff import abcd as abcd_mod
    from xanthos_b200.routing import mrtm as mrtm_mod

    torch.cuda.set_device(local_rank)
    C.lib()
    ncell, nmonths = args.ncell, args.nmonths
    world, pm, ab, end_yr = build_inputs(member_seed=1 + rank, ncell=ncell, nmonths=nmonths)
    settings = world.settings()
    ndays = month_days_mod4(nmonths, START_YR)
    spin_ro, spin_rt = min(RUNOFF_SPINUP, nmonths), min(ROUTING_SPINUP, nmonths)
    lc_years = pm['lc_years']

    # ---- static, device-resident -------------------------------------------------------------------
    dsid = mrtm_mod.downstream(world.coords, world.flow_dir, settings)
    upid = mrtm_mod.upstream(world.coords, dsid, settings)
    um = mrtm_mod.upstream_genmatrix(upid)
    rows = abcd_mod._basin_rows(world.n_basins, world.basin_ids, world.n_basins)
    plan = abcd_mod.basin_plan(rows, world.n_basins)
    d_L, d_V, d_A = C.dev_vector(world.flow_dist), C.dev_vector(world.velocity), C.dev_vector(world.area)
    d_Akm3 = d_A * 1e-6
    d_pars = torch.from_numpy(ab['pars']).cuda()

    # ---- host inputs in pinned memory (what a DataLoader would hold) --------------------------------
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
        return t.numpy()
    forc_names = ('tair_load', 'TMIN_load', 'rhs_load', 'wind_load', 'rsds_load', 'rlds_load')
    host = {k: pinned(pm[k]) for k in forc_names}
    host['precip'] = pinned(ab['precip'])
    host['tmin'] = pinned(ab['tmin'])
    host['lct_load'] = pinned(pm['lct_load'])
    tables = {k: pm[k] for k in pm if k not in forc_names + ('lct_load', 'tairprev_load')}
    h2d_bytes = sum(v.nbytes for v in host.values())

    # ---- device-resident copies for the `value` leg ---------------------------------------------------
    dev = {k: C.Field.from_host(host[k]) for k in forc_names + ('precip', 'tmin')}
    # further members of the step: different forcing (the cells rolled by 7), the same static data
    devs = [dev] + [{k: C.Field.from_host(np.roll(host[k], 7 * j, axis=0)) for k in forc_names + ('precip', 'tmin')}
                    for j in range(1, MEMBERS_PER_STEP)]
    d_lct = pm_mod.stage_land_cover(host['lct_load'], dev['tair_load'].ld)
    d_elev = C.dev_vector(pm['elev'])
    torch.cuda.synchronize()

    def data_ns(src, lct):
        d = dict(tables)
        d.update({k: src[k] for k in forc_names})
        d['lct_load'] = lct
        d['elev'] = d_elev
        return SimpleNamespace(**d)

    stage_ms = {'pm': [], 'abcd': [], 'mrtm': [], 'agg': []}
    stage_ms_single = {'pm': [], 'abcd': [], 'mrtm': [], 'agg': []}

    def device_step(record=False, nm=MEMBERS_PER_STEP):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        pets = [pm_mod.run_pmpet_device(data_ns(d, d_lct), ncell, NLCS, START_YR, end_yr, pm['water_idx'],
                                        pm['snow_idx'], lc_years) for d in devs[:nm]]
        ev[1].record()
        ress = [abcd_mod.run_device(plan, d_pars, pet, d['precip'], d['tmin'], nmonths, spin_ro)
                for pet, d in zip(pets, devs[:nm])]
        ev[2].record()
        if nm == 1:
            routed = [mrtm_mod.route_device(um, ress[0]['q'], d_L, d_V, d_A, ndays, DT, spin_rt)]
        else:
            routed = mrtm_mod.route_device_batch(um, [r['q'] for r in ress], d_L, d_V, d_A, ndays, DT, spin_rt)
        ev[3].record()
        # basin aggregates: runoff in km3/month and mean streamflow, [members, 2, nmonths, n_basins]
        agg = torch.empty((nm, 2, nmonths, world.n_basins), dtype=torch.float64, device='cuda')
        for j, (res, (chs, avg, inst)) in enumerate(zip(ress, routed)):
            C.check(C.lib().xan_basin_sum(plan._plan, C.ptr(res['q'].t), C.ptr(d_Akm3), nmonths, res['q'].ld,
                                          C.ptr(agg[j, 0]), C.stream_ptr()))
            C.check(C.lib().xan_basin_sum(plan._plan, C.ptr(avg.t), None, nmonths, avg.ld, C.ptr(agg[j, 1]),
                                          C.stream_ptr()))
        if world_size > 1:
            gathered = [torch.empty_like(agg) for _ in range(world_size)]
            dist.all_gather(gathered, agg)
        ev[4].record()
        if record:
            torch.cuda.synchronize()
            for k, i in (('pm', 0), ('abcd', 1), ('mrtm', 2), ('agg', 3)):
                (stage_ms if nm == MEMBERS_PER_STEP else stage_ms_single)[k].append(ev[i].elapsed_time(ev[i + 1]))
        return pets[0], ress[0], routed[0][0], routed[0][1], agg

    d2h_bytes_holder = [0]

    timeline = []

    def e2e_step():
        """Reference-facing plug-in calls on host buffers (the Components.simulation sequence)."""
        tl = os.environ.get('XANTHOS_BENCH_TIMELINE')
        marks = []

        def mark(name):
            if tl:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e, time.perf_counter()))
        C.forget_all()
        mark('start')
        with C.async_host():          # exactly what Components.simulation does (xanthos_b200/components.py)
            pet = pm_mod.run_pmpet(data_ns(host, host['lct_load']), ncell, NLCS, START_YR, end_yr, pm['water_idx'],
                                   pm['snow_idx'], lc_years)
            C.prefetch(host['precip'])   # queued behind the PM forcing: PM runs under these two uploads
            C.prefetch(host['tmin'])
            mark('pm')
            pet, aet, q, sav = abcd_mod.abcd_execute(n_basins=world.n_basins, basin_ids=world.basin_ids, pet=pet,
                                                     precip=host['precip'], tmin=host['tmin'], calib_file=ab['pars'],
                                                     n_months=nmonths, spinup_steps=spin_ro, jobs=-1)
            mark('abcd')
            chs, avg, inst = mrtm_mod.route(um, q, world.flow_dist, world.velocity, world.area, ndays, DT, spin_rt)
            mark('mrtm')
        mark('d2h_done')
        d2h_bytes_holder[0] = sum(a.nbytes for a in (pet, aet, q, sav, chs, avg, inst))
        if tl:
            torch.cuda.synchronize()
            timeline.append({n: (round(marks[0][1].elapsed_time(e), 2), round((t - marks[0][2]) * 1e3, 2))
                             for n, e, t in marks[1:]})
        return float(avg[0, -1]) + float(pet[0, 0]) + float(sav[-1, -1])

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall], dtype=torch.float64, device='cuda')
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    # ---- measure ----------------------------------------------------------------------------------------
    import gc
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = None
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler.mark()
    gc.collect()
    gc.disable()          # a generation-2 collection while the first steps are being queued leaves the GPU idle
    barrier()
    launches0 = C.launch_count
    try:
        dev_ms, _ = timed(device_step, args.steps, 0)
    finally:
        gc.enable()
    launches = (C.launch_count - launches0) // max(args.steps, 1)
    clocks = sampler.stop()
    for _ in range(3):
        device_step(record=True)
    # the one-member step (round 1's definition of `value`), for continuity
    single_ms = None
    if MEMBERS_PER_STEP > 1:
        single_ms, _ = timed(lambda: device_step(nm=1), max(3, args.steps // 2), 2)
        single_ms /= max(3, args.steps // 2)
        for _ in range(3):
            device_step(record=True, nm=1)
    gc.collect()
    gc.disable()          # a generation-2 collection in the middle of a 100 ms step is a 50 ms hiccup
    e2e_steps = max(3, min(args.steps, 9))
    e2e_each = []
    try:
        for _ in range(3):
            e2e_step()
        for _ in range(e2e_steps):          # every step is bracketed on its own: its result is read on the host
            barrier()
            t0 = time.perf_counter()
            e2e_step()
            barrier()
            e2e_each.append((time.perf_counter() - t0) * 1e3)
    finally:
        gc.enable()
    t = torch.tensor(e2e_each, dtype=torch.float64, device='cuda')
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_each = [float(v) for v in t.cpu()]
    # The GPU boxes are shared hosts: single steps are sometimes stretched by 80 - 700 ms of host-side stalls
    # (seen as a GPU waiting idle for the next launch).  The median step is reported; the mean and every step are kept.
    e2e_single_ms = float(np.median(e2e_each))
    C.forget_all()

    # ---- end to end, the way a multi-scenario user runs it: xanthos_b200.ensemble.run_ensemble - every member's forcing
    # comes from pinned host memory and its requested outputs (q, avgchflow) + basin aggregates go back to the host
    # inside the timing; the copies of neighbouring members overlap the kernels (h2d / compute / d2h streams)
    from xanthos_b200 import ensemble as ens
    statics = ens.EnsembleStatics(ncell, START_YR, end_yr, tables, host['lct_load'], pm['elev'], pm['water_idx'],
                                  pm['snow_idx'], lc_years, NLCS, world.n_basins, world.basin_ids, ab['pars'], world.area,
                                  world.flow_dist, world.velocity, um, ndays, DT, spin_ro, spin_rt)
    member_a = {k: host[k] for k in ens.FORCING}
    member_b = {k: pinned(np.roll(host[k], 7, axis=0)) for k in ens.FORCING}    # a second, different member
    n_mem = max(6, min(args.steps, 24))
    sink = []

    def ensemble_ms(ma, mb, reps=3):
        """Wall time of run_ensemble over n_mem members per rank: the median of `reps` runs (max over ranks each) - the GPU
        boxes are shared hosts, and once in a while the first uploads of a run crawl at a tenth of the link speed."""
        members = [ma if k % 2 == 0 else mb for k in range(n_mem * world_size)]
        ens.run_ensemble(statics, members[:4 * world_size], on_result=lambda i, r: sink.append(float(r['avgchflow'][0, -1])))
        each, stats = [], None
        for _ in range(reps):
            gc.collect()
            gc.disable()
            try:
                barrier()
                t0 = time.perf_counter()
                er = ens.run_ensemble(statics, members, on_result=lambda i, r: sink.append(float(r['avgchflow'][0, -1])))
                barrier()
                ms = (time.perf_counter() - t0) * 1e3
            finally:
                gc.enable()
            if rank == 0 and 'timeline' in er['stats']:
                print('ensemble timeline (ms since the first upload): %s\nhost pool: %s' % (
                    json.dumps(er['stats'].pop('timeline')), C.host_pool.stats()), file=sys.stderr, flush=True)
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            if world_size > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            each.append(float(t[0]))
            stats = er['stats']
        stats = dict(stats, ms_each_run=[round(v, 2) for v in each])
        return float(np.median(each)), stats
    ens64_ms, ens64_stats = ensemble_ms(member_a, member_b)                      # float64 arrays across the link
    # lossless single-precision transport: what a loader does once per member when the values allow it
    ens_ms, ens_stats = ensemble_ms(ens.lossless_float32(member_a), ens.lossless_float32(member_b), reps=5)
    del member_b
    e2e_ms = ens_ms / n_mem

    # ---- calibration objective (BASELINE.json metric, second half): one differential-evolution generation =
    # 64 candidate parameter sets x every basin, each a full spin-up + simulation + basin sum + KGE distance
    calib = None
    if not args.no_calib:
        from xanthos_b200.calibrate import calibrate_abcd as cal
        pet_f = device_step()[0]
        ev = cal.BasinEvaluator(world.basin_ids, world.area, dev['precip'], pet_f, dev['tmin'], nmonths, spin_ro,
                                'km3_per_mth')
        rng = np.random.default_rng(4)
        P = CALIB_POP
        bnums = np.arange(1, world.n_basins + 1)
        lo = np.array([b[0] for b in cal.BOUNDS_SNOW])
        hi = np.array([b[1] for b in cal.BOUNDS_SNOW])
        cpars = lo + (hi - lo) * rng.random((world.n_basins, P, 5))
        _, series = ev.evaluate(bnums, np.broadcast_to(ab['pars'][:, None, :], (world.n_basins, 1, 5)).copy(),
                                np.ones((world.n_basins, nmonths)), want_series=True)
        obs = series[:, 0, :] * (1 + rng.normal(0, 0.05, (world.n_basins, nmonths)))   # "VIC-like" observations
        cal_steps = max(1, min(args.steps, 3))
        cal_ms, cal_wall = timed(lambda: ev.evaluate(bnums, cpars, obs), cal_steps, 1)
        n_eval = world_size * world.n_basins * P
        calib = {'value': n_eval / (max(cal_ms, cal_wall) / cal_steps * 1e-3), 'unit': 'param-sets/s',
                 'ms_per_generation': max(cal_ms, cal_wall) / cal_steps, 'population': P, 'basins': int(world.n_basins),
                 'months': nmonths, 'spinup': spin_ro,
                 'cell_month_steps_per_s': n_eval / world.n_basins * ncell * (nmonths + spin_ro)
                 / (max(cal_ms, cal_wall) / cal_steps * 1e-3),
                 'note': 'xan_abcd_kge_batch through BasinEvaluator.evaluate: parameters and observations H2D, '
                         'distances D2H inside the timing; forcing resident'}
        # whole differential-evolution loop (objective + generation logic), device driver against host driver
        gens = 12
        robs_de = obs

        def de_loop(driver):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if driver == 'device':
                r = cal.differential_evolution_device(ev, bnums, robs_de, cal.BOUNDS_SNOW, popsize=13, maxiter=gens,
                                                      tol=0.0, seed=4)
            else:
                r = cal.differential_evolution_batched(lambda x, idx: ev.evaluate(bnums[idx], x, robs_de[idx]),
                                                       len(bnums), cal.BOUNDS_SNOW, popsize=13, maxiter=gens, tol=0.0,
                                                       seed=4)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0), int(r['nfev'].sum())
        de_loop('device')
        t_dev, nf_dev = de_loop('device')
        t_host, nf_host = de_loop('host')
        calib['de_loop'] = {'generations': gens, 'population': 65,
                            'device_driver_param_sets_per_s': world_size * nf_dev / t_dev,
                            'host_driver_param_sets_per_s': world_size * nf_host / t_host,
                            'note': 'Latin-hypercube initialisation + %d generations for all basins; device driver = '
                                    'xan_de_trial / xan_abcd_kge_batch / xan_de_select without host round trips' % gens}
        del ev, pet_f

    # ---- post-processing scans on the resident runoff: drought thresholds + statistics,
    # basin aggregation.  Device time per call, algorithmic bytes, and the numpy port beside it (rank 0, N = 1).
    postproc = None
    if not args.no_calib:
        from xanthos_b200.drought import drought_stats as dr
        from xanthos_b200.diagnostics import time_series as tsm
        qf = device_step()[1]['q']                                  # Field [nmonths][ld]
        thr = dr.getthresh_device(qf.t, ncell, 12)

        def ev_ms(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        ms_thr = ev_ms(lambda: dr.getthresh_device(qf.t, ncell, 12))
        ms_sid = ev_ms(lambda: dr.droughtstats_device(qf.t, ncell, thr))
        ms_grp = ev_ms(lambda: tsm.group_sum_device(world.basin_ids, qf.t))
        cmf = float(ncell) * nmonths
        postproc = {'drought_stats': {'ms': ms_sid, 'algorithmic_bytes': 32 * cmf,
                                      'achieved_gbs': 32 * cmf / (ms_sid * 1e-3) / 1e9},
                    'drought_thresholds': {'ms': ms_thr, 'algorithmic_bytes': 8 * cmf + 8 * 12 * ncell,
                                           'achieved_gbs': (8 * cmf + 96 * ncell) / (ms_thr * 1e-3) / 1e9},
                    'basin_group_sum': {'ms': ms_grp, 'algorithmic_bytes': 8 * cmf,
                                        'achieved_gbs': 8 * cmf / (ms_grp * 1e-3) / 1e9,
                                        'note': 'group plan (cells sorted by basin) cached on the device'},
                    'cell_months_per_s': cmf / ((ms_thr + ms_sid + ms_grp) * 1e-3)}
        if world_size == 1 and not args.no_cpu_baseline:
            from oracle import postproc as opp
            qh = qf.t[:, :ncell].cpu().numpy()                      # [ntime, ngrid]
            t0 = time.perf_counter()
            thr_h = opp.getthresh(qh, 12)
            opp.droughtstats(qh, thr_h)
            t1 = time.perf_counter()
            sub = slice(0, min(ncell, 4000))                        # the reference's double loop is O(ncell x ntime) Python
            opp.aggregation_map(np.asarray(world.basin_ids)[sub], np.ascontiguousarray(qh.T[sub]))
            t2 = time.perf_counter()
            t_cpu = (t1 - t0) + (t2 - t1) * ncell / (sub.stop - sub.start)
            postproc['cpu_port'] = {'cell_months_per_s': cmf / t_cpu, 'cores': 1,
                                    'sample': 'thresholds + statistics on all cells; aggregation on %d cells, '
                                              'extrapolated linearly' % (sub.stop - sub.start)}
        del qf, thr

    cm = float(ncell) * nmonths
    ms_step = dev_ms / args.steps
    value = world_size * MEMBERS_PER_STEP * cm / (ms_step * 1e-3)
    e2e_val = world_size * cm / (e2e_ms * 1e-3)

    if rank != 0:
        return
    if timeline:
        print('e2e timeline, ms since step start as (device event on the compute stream, host wall): %s'
              % json.dumps(timeline), file=sys.stderr, flush=True)
        print('host pool: %s' % C.host_pool.stats(), file=sys.stderr, flush=True)
    peaks, peak_src = _peaks()
    med = {k: float(np.median(v)) for k, v in stage_ms.items()}
    # per launch: PM and ABCD run once per member (their time is the sum over the step's members), routing once per step
    nmem = MEMBERS_PER_STEP
    per_kernel = {
        'pm_pet_kernel': dict(ms=med['pm'] / nmem, alg_bytes=BYTES_PM * cm),
        'abcd_spinup+reinit+sim': dict(ms=med['abcd'] / nmem, alg_bytes=BYTES_ABCD * cm),
        'mrtm_warp_kernel': dict(ms=med['mrtm'], alg_bytes=BYTES_MRTM * cm * nmem),
    }
    for v in per_kernel.values():
        v['gbs'] = v['alg_bytes'] / (v['ms'] * 1e-3) / 1e9
        v['frac_hbm'] = v['gbs'] / peaks['hbm_gbs']
    # which forest kernel XAN_MRTM_AUTO ran (csrc/mrtm.cu AUTO_DEFAULT_SKEW; XANTHOS_MRTM_AUTO=tree|skew overrides)
    if os.environ.get('XANTHOS_MRTM_AUTO', 'skew') == 'tree':
        mrtm_kernel = 'mrtm_warp_kernel<1,640>'
    elif MEMBERS_PER_STEP > 1 and os.environ.get('XANTHOS_MRTM_SKEW_MEMBERS', '2') != '1':
        mrtm_kernel = 'mrtm_skew_kernel<%s,2,128>' % os.environ.get('XANTHOS_MRTM_SKEW_KM', '4')   # <cells per lane, members per launch, threads>
    else:
        mrtm_kernel = 'mrtm_skew_kernel<%s,1,256>' % os.environ.get('XANTHOS_MRTM_SKEW_K', '2')
    # DRAM traffic per launch and pipe utilisation from the committed ncu --set full capture of the same workload
    # (profiles/r02b_kernels.json, written by tools/ncu_summary.py from the .ncu-rep of this round)
    traffic, ncu_k, ncu_src = {}, {}, None
    for fname in ('r02b_kernels.json', 'r02_kernels.json', 'r01c_traffic.json'):
        try:
            with open(os.path.join(ROOT, 'profiles', fname)) as f:
                tk = json.load(f)['kernels']
            find = lambda prefix: next((v for k, v in tk.items() if k.startswith(prefix)), None)       # noqa: E731
            pmk, spk, smk = find('pm_pet_fast_kernel'), find('abcd_spinup_kernel'), find('abcd_sim_kernel')
            mrk = find(mrtm_kernel) or find(mrtm_kernel.split('<')[0])
            if not (pmk and spk and smk and mrk):
                continue
            ncu_k = {'pm_pet_kernel': pmk, 'abcd_spinup+reinit+sim': smk, 'mrtm_warp_kernel': mrk}
            if (ncell, nmonths) == (NCELL, NMONTHS):
                traffic = {'pm_pet_kernel': pmk['dram_bytes_per_launch'],
                           'abcd_spinup+reinit+sim': spk['dram_bytes_per_launch'] + smk['dram_bytes_per_launch'],
                           'mrtm_warp_kernel': mrk['dram_bytes_per_launch']}
            ncu_src = 'profiles/' + fname
            break
        except Exception:
            continue
    dom = max(per_kernel, key=lambda k: per_kernel[k]['ms'])
    # The dominant kernel is not bandwidth-bound: 2 x sum(nt) strictly sequential sub-steps, each a dependent chain
    # (S -> F -> exchange -> row sum -> S).  Its floor is sub-steps x the chain of an ISOLATED warp, measured by
    # tools/microbench (profiles/r02_fp64_peak.json: mrtm_chain_cycles for row lengths 1..9).
    latency_model, fp64_peaks = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_fp64_peak.json')) as f:
            mb = json.load(f)
        nsub = int(sum(int(d) * 24 * 3600 // DT for d in ndays[:spin_rt])) + int(sum(int(d) * 24 * 3600 // DT for d in ndays))
        chain = mb['mrtm_chain_cycles_nt1_9'][3]          # rows are padded to the longest row of a warp: 4 - 6 terms
        mhz = clocks['sm_mhz'] or mb['sm_mhz']
        floor_ms = nsub * chain / (mhz * 1e3)
        latency_model = {'sub_steps': nsub, 'chain_cycles_isolated_warp_nt4': chain, 'sm_mhz': mhz, 'floor_ms': floor_ms,
                         'measured_ms': per_kernel['mrtm_warp_kernel']['ms'],
                         'members_per_launch': nmem,
                         'measured_ms_per_member': per_kernel['mrtm_warp_kernel']['ms'] / nmem,
                         'frac_of_floor': floor_ms / (per_kernel['mrtm_warp_kernel']['ms'] / nmem),
                         'us_per_sub_step': per_kernel['mrtm_warp_kernel']['ms'] / nmem * 1e3 / nsub,
                         'source': 'profiles/r02_fp64_peak.json (tools/microbench/fp64_peak.cu)',
                         'note': 'the floor is the dependent chain of one isolated warp of the warp-dataflow kernel (row of 4 '
                                 'terms); the warps of an SM share its issue ports (DESIGN.md section 4)'}
        fp64_peaks = {'dfma_tflops': mb['dfma_tflops'], 'dmul_dadd_tflops': mb['dmul_dadd_tflops'],
                      'source': 'profiles/r02_fp64_peak.json'}
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': per_kernel[dom]['gbs'], 'peak': peaks['hbm_gbs'],
                'unit': 'GB/s', 'frac': per_kernel[dom]['frac_hbm'], 'traffic': traffic.get(dom),
                'traffic_source': '%s (ncu dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch)' % ncu_src,
                'algorithmic_bytes': per_kernel[dom]['alg_bytes'], 'peak_source': peak_src,
                'share_of_step': med['mrtm'] / (med['pm'] + med['abcd'] + med['mrtm']),
                'mrtm_kernel': mrtm_kernel,
                'limiter': 'latency of the sequential sub-step recurrence and issue slots - NOT HBM; "bound": "hbm" only '
                           'names the peak the contract asks to report against',
                'latency_model': latency_model, 'fp64_peaks': fp64_peaks,
                'note': 'dominant kernel is bound by the latency of its sequential sub-step recurrence and by issue slots, not by HBM; see DESIGN.md',
                'kernels': {k: {'ms': round(v['ms'], 4), 'achieved_gbs': round(v['gbs'], 2),
                                'frac_hbm': round(v['frac_hbm'], 5), 'algorithmic_bytes': v['alg_bytes'],
                                'traffic': traffic.get(k),
                                'fp64_pipe_busy_pct': (ncu_k.get(k) or {}).get('fp64_pipe_pct'),
                                'issue_active_pct': (ncu_k.get(k) or {}).get('issue_pct')} for k, v in per_kernel.items()}}
    line = {
        'metric': 'cell-months/s (PM+ABCD+MRTM)', 'value': value, 'unit': 'cell-months/s', 'n_gpus': world_size,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': bench_config(ncell, nmonths),
        'mrtm_plan': um.info,
        'e2e': {'value': e2e_val, 'unit': 'cell-months/s',
                'h2d_bytes_per_step': int(ens_stats['h2d_bytes'] // max(ens_stats['members_local'], 1)),
                'd2h_bytes_per_step': int(ens_stats['d2h_bytes'] // max(ens_stats['members_local'], 1)),
                'ms_per_step': e2e_ms, 'steps': n_mem,
                'statistic': 'median of %d runs of %d members' % (len(ens_stats['ms_each_run']), n_mem),
                'ms_each_run': ens_stats['ms_each_run'],
                'mode': 'xanthos_b200.ensemble.run_ensemble: %d members per GPU back to back (two per routing launch), forcing '
                        'from pinned host memory, outputs q + avgchflow + basin aggregates to the host, copies of '
                        'neighbouring members overlapped with the kernels; the forcing values are single precision and cross the link as float32 '
                        '(ensemble.lossless_float32: verified exact per array, results bit-identical)' % n_mem,
                'float64_transport': {'ms_per_step': ens64_ms / n_mem, 'ms_each_run': ens64_stats['ms_each_run'], 'value': world_size * cm / (ens64_ms / n_mem * 1e-3),
                                      'h2d_bytes_per_step': int(ens64_stats['h2d_bytes'] // max(ens64_stats['members_local'], 1))},
                'single_member': {'ms_per_step': e2e_single_ms, 'value': world_size * cm / (e2e_single_ms * 1e-3),
                                  'h2d_bytes_per_step': int(h2d_bytes), 'd2h_bytes_per_step': int(d2h_bytes_holder[0]),
                                  'statistic': 'median step', 'ms_each_step': [round(v, 2) for v in e2e_each],
                                  'note': 'run_pmpet / abcd_execute / route on host arrays, all six outputs back to the '
                                          'host, nothing overlapped across members (round-1 definition of e2e)'}},
        'single_member_step': None if single_ms is None else {
            'ms_per_step': single_ms, 'value': world_size * cm / (single_ms * 1e-3),
            'stages_ms': {k: round(float(np.median(v)), 4) for k, v in stage_ms_single.items() if v},
            'note': 'one member per step and per routing launch (round 1 definition of `value`)'},
        'gpu_launches': int(launches) * args.steps,
        'gpu_launches_per_step': int(launches),
        'clocks': clocks,
        'roofline': roofline,
    }
    if calib is not None:
        line['calib'] = calib
    if postproc is not None:
        for k in ('drought_stats', 'drought_thresholds', 'basin_group_sum'):
            postproc[k]['frac_hbm'] = postproc[k]['achieved_gbs'] / peaks['hbm_gbs']
        line['postproc'] = postproc
    if world_size == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(world, pm, ab, nmonths)
        if calib is not None:
            line['cpu_baseline']['calib'] = cpu_calib_baseline(world, pm, ab)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (`--workload`): one JSON line each, same keys as the main line
# ------------------------------------------------------------------------------------------------
def _line(metric, unit, value, ms_per_step, args, world_size, config, extra):
    line = {'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world_size, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': config.pop('scaling', 'weak'),
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config}
    line.update(extra)
    return line


def run_workload(args, rank, world_size, local_rank):
    import torch
    import torch.distributed as dist
    from xanthos_b200 import synthetic, _cuda as C
    from xanthos_b200.pet import hargreaves_samani as hs_mod
    from xanthos_b200.runoff import abcd as abcd_mod
    from xanthos_b200.routing import mrtm as mrtm_mod
    torch.cuda.set_device(local_rank)
    C.lib()
    peaks, peak_src = _peaks()
    world = synthetic.make_world(seed=0)
    ncell = world.ncell
    name = args.workload

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3], dtype=torch.float64, device='cuda')
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / steps, float(t[1]) / steps

    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory().numpy()

    def pinned_any(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    sampler = ClockSampler(local_rank)
    sampler.start()
    line = None
    if name == 'hs_abcd_1140':
        # BASELINE config 2: Hargreaves-Samani PET + ABCD runoff, no routing, 67,420 cells x 1,140 months (2006-2100)
        sy, ey, m, spin = 2006, 2100, 1140, 360
        hs = synthetic.hs_inputs(world, sy, ey, seed=2 + rank)
        ab = synthetic.abcd_inputs(world, m, seed=1 + rank, with_pet=False)
        host = {k: pinned(np.nan_to_num(hs[k])) for k in ('hs_tas', 'hs_tmax', 'hs_tmin')}
        host['precip'], host['tmin'] = pinned(ab['precip']), pinned(np.nan_to_num(ab['tmin']))
        dev = {k: C.Field.from_host(v) for k, v in host.items()}
        rows = abcd_mod._basin_rows(world.n_basins, world.basin_ids, world.n_basins)
        plan = abcd_mod.basin_plan(rows, world.n_basins)
        d_pars = torch.from_numpy(ab['pars']).cuda()
        cfg = SimpleNamespace(ncell=ncell, nmonths=m, StartYear=sy, EndYear=ey)

        def dev_step():
            pet = hs_mod.execute_device(dev['hs_tas'], dev['hs_tmax'], dev['hs_tmin'], world.coords[:, 2], sy)
            return abcd_mod.run_device(plan, d_pars, pet, dev['precip'], dev['tmin'], m, spin)

        def e2e_step():
            C.forget_all()
            with C.async_host():
                data = SimpleNamespace(coords=world.coords, hs_tas=host['hs_tas'], hs_tmax=host['hs_tmax'],
                                       hs_tmin=host['hs_tmin'])
                pet = hs_mod.execute(cfg, data)
                out = abcd_mod.abcd_execute(world.n_basins, world.basin_ids, pet, host['precip'], host['tmin'], ab['pars'],
                                            m, spin, -1)
            return float(out[2][0, -1])
        sampler.mark()
        ms, _ = timed(dev_step, args.steps, args.warmup)
        clocks = sampler.stop()
        _, e2e_ms = timed(e2e_step, max(2, min(args.steps, 4)), 1)
        cm = float(ncell) * m
        alg = (32 + 48 + 24 * spin / m) * cm
        line = _line('cell-months/s (HS+ABCD)', 'cell-months/s', world_size * cm / (ms * 1e-3), ms, args, world_size,
                     {'workload': name, 'ncell': ncell, 'nmonths': m, 'runoff_spinup': spin,
                      'parallelism': 'member-per-gpu', 'l2': 'inputs (5 x 615 MB) exceed the 126 MB L2'},
                     {'e2e': {'value': world_size * cm / (e2e_ms * 1e-3), 'unit': 'cell-months/s', 'ms_per_step': e2e_ms,
                              'h2d_bytes_per_step': int(sum(v.nbytes for v in host.values())),
                              'd2h_bytes_per_step': int(5 * ncell * m * 8)},
                      'roofline': {'bound': 'hbm', 'kernel': 'hs_pet + abcd_spinup + abcd_sim', 'achieved': alg / (ms * 1e-3) / 1e9,
                                   'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': alg / (ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                                   'algorithmic_bytes': alg, 'traffic': None, 'peak_source': peak_src},
                      'clocks': clocks, 'gpu_launches': 5 * args.steps})
    elif name == 'mrtm_hourly_360':
        # BASELINE config 3: routing only, hourly sub-steps over 360 months (+ 360 months of spin-up)
        m, spin, dt = 360, 360, 3600
        q_h = pinned(synthetic.runoff_input(world, m, seed=3 + rank))
        q = C.Field.from_host(q_h)
        s = world.settings()
        um = mrtm_mod.upstream_genmatrix(mrtm_mod.upstream(world.coords, mrtm_mod.downstream(world.coords, world.flow_dir, s), s))
        L, V, A = C.dev_vector(world.flow_dist), C.dev_vector(world.velocity), C.dev_vector(world.area)
        nd = month_days_mod4(m, START_YR)
        sampler.mark()
        ms, _ = timed(lambda: mrtm_mod.route_device(um, q, L, V, A, nd, dt, spin), args.steps, args.warmup)
        clocks = sampler.stop()

        def e2e_step():
            C.forget_all()
            return float(mrtm_mod.route(um, q_h, world.flow_dist, world.velocity, world.area, nd, dt, spin)[1][0, -1])
        _, e2e_ms = timed(e2e_step, max(2, min(args.steps, 4)), 1)
        cm = float(ncell) * m
        nsub = int(sum(int(d) * 24 for d in nd)) * 2
        alg = BYTES_MRTM * cm
        line = _line('cell-months/s (MRTM hourly)', 'cell-months/s', world_size * cm / (ms * 1e-3), ms, args, world_size,
                     {'workload': name, 'ncell': ncell, 'nmonths': m, 'routing_spinup': spin, 'dt_s': dt,
                      'sub_steps': nsub, 'parallelism': 'member-per-gpu', 'mrtm_plan': um.info},
                     {'e2e': {'value': world_size * cm / (e2e_ms * 1e-3), 'unit': 'cell-months/s', 'ms_per_step': e2e_ms,
                              'h2d_bytes_per_step': int(q_h.nbytes), 'd2h_bytes_per_step': int(2 * q_h.nbytes + 8 * ncell)},
                      'roofline': {'bound': 'hbm', 'kernel': 'mrtm_warp_kernel', 'achieved': alg / (ms * 1e-3) / 1e9,
                                   'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': alg / (ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                                   'algorithmic_bytes': alg, 'traffic': None, 'peak_source': peak_src,
                                   'us_per_sub_step': ms * 1e3 / nsub,
                                   'note': 'bound by the sequential sub-step recurrence, not by HBM (DESIGN.md section 4)'},
                      'clocks': clocks, 'gpu_launches': args.steps})
    elif name == 'calib_235x64x200':
        # BASELINE config 4: 235 basins x 64-member differential evolution x 200 generations; basins sharded over the
        # ranks by cell count (strong scaling: the total work is fixed)
        from xanthos_b200.calibrate import calibrate_abcd as cal
        from xanthos_b200 import sharding
        m, spin, P, gens = 360, 360, 64, 200
        ab = synthetic.abcd_inputs(world, m, seed=1)
        tmin = np.nan_to_num(ab['tmin'])
        ev = cal.BasinEvaluator(world.basin_ids, world.area, ab['precip'], ab['pet'], tmin, m, spin, 'km3_per_mth')
        all_b = np.arange(1, world.n_basins + 1)
        _, series = ev.evaluate(all_b, np.broadcast_to(ab['pars'][:, None, :], (world.n_basins, 1, 5)).copy(),
                                np.ones((world.n_basins, m)), want_series=True)
        obs = series[:, 0, :] * (1 + np.random.default_rng(4).normal(0, 0.05, (world.n_basins, m)))   # "VIC-like"
        mine = np.asarray(sharding.partition_basins(world.basin_ids, world_size)[rank])
        cells = np.bincount(np.asarray(world.basin_ids).astype(int), minlength=world.n_basins + 1)

        def loop():
            return cal.differential_evolution_device(ev, mine, obs[mine - 1], cal.BOUNDS_SNOW, maxiter=gens, tol=0.0,
                                                     seed=4, pop_members=P)
        loop() if args.warmup and gens <= 20 else cal.differential_evolution_device(
            ev, mine, obs[mine - 1], cal.BOUNDS_SNOW, maxiter=3, tol=0.0, seed=4, pop_members=P)
        sampler.mark()
        barrier()
        t0 = time.perf_counter()
        r = loop()
        barrier()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device='cuda')
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        clocks = sampler.stop()
        ms = float(t[0])
        n_eval = world.n_basins * P * (gens + 1)
        best = 1 - r['fun']
        line = _line('calibration param-sets/s', 'param-sets/s', n_eval / (ms * 1e-3), ms, args, world_size,
                     {'workload': name, 'basins': int(world.n_basins), 'population': P, 'generations': gens, 'months': m,
                      'spinup': spin, 'scaling': 'strong',
                      'parallelism': 'basins over %d rank(s), LPT by cell count; rank 0 holds %d basins / %d cells, largest '
                                     'basin %d cells' % (world_size, len(mine), int(cells[mine].sum()), int(cells.max()))},
                     {'e2e': {'value': n_eval / (ms * 1e-3), 'unit': 'param-sets/s', 'ms_per_step': ms,
                              'h2d_bytes_per_step': int(obs[mine - 1].nbytes),
                              'd2h_bytes_per_step': int(len(mine) * P * 6 * 8),
                              'note': 'the whole DE loop is one public call; forcing staged once before the loop'},
                      'roofline': {'bound': 'hbm', 'kernel': 'kge_pop_pass_kernel', 'achieved': None, 'peak': peaks['hbm_gbs'],
                                   'unit': 'GB/s', 'frac': None, 'traffic': None,
                                   'note': 'FP64-pipe bound (profiles/r01c_calib_kernels.csv), see roofline of the main line'},
                      'kge_rank0': {'median': float(np.median(best)), 'min': float(best.min())},
                      'ms_per_generation': ms / (gens + 1), 'clocks': clocks, 'gpu_launches': 6 * (gens + 1)})
    elif name == 'ensemble_64x1032':
        # BASELINE config 5: 64-member ensemble, full PM + ABCD + MRTM 2015-2100 (1,032 months), members over the ranks
        from xanthos_b200 import ensemble as ens
        sy, ey = 2015, 2100
        m = (ey - sy + 1) * 12
        n_members = int(os.environ.get('XANTHOS_BENCH_MEMBERS', '64'))
        pm = synthetic.pm_inputs(world, sy, ey, nlcs=NLCS, seed=1 + rank)
        ab = synthetic.abcd_inputs(world, m, seed=1 + rank, with_pet=False)
        tables = {k: v for k, v in pm.items() if k not in ens.PM_FORCING + ('lct_load', 'tairprev_load')}
        f32 = os.environ.get('XANTHOS_BENCH_F64_TRANSPORT') is None     # single-precision VALUES, float32 across the link
        rnd = (lambda a: a.astype(np.float32)) if f32 else (lambda a: a.astype(np.float32).astype(np.float64))
        mem_a = {k: pinned_any(rnd(np.nan_to_num(pm[k]))) for k in ens.PM_FORCING}
        mem_a['precip'], mem_a['tmin'] = pinned_any(rnd(ab['precip'])), pinned_any(rnd(np.nan_to_num(ab['tmin'])))
        mem_b = {k: pinned_any(np.roll(v, 11, axis=0)) for k, v in mem_a.items()}
        s = world.settings()
        um = mrtm_mod.upstream_genmatrix(mrtm_mod.upstream(world.coords, mrtm_mod.downstream(world.coords, world.flow_dir, s), s))
        nd = month_days_mod4(m, sy)
        st = ens.EnsembleStatics(ncell, sy, ey, tables, pm['lct_load'], pm['elev'], pm['water_idx'], pm['snow_idx'],
                                 pm['lc_years'], NLCS, world.n_basins, world.basin_ids, ab['pars'], world.area,
                                 world.flow_dist, world.velocity, um, nd, DT, RUNOFF_SPINUP, ROUTING_SPINUP)
        members = [mem_a if k % 2 == 0 else mem_b for k in range(n_members)]
        sink = []
        ens.run_ensemble(st, members[:4 * world_size], on_result=lambda i, r: sink.append(float(r['avgchflow'][0, -1])))
        sampler.mark()
        barrier()
        t0 = time.perf_counter()
        er = ens.run_ensemble(st, members, on_result=lambda i, r: sink.append(float(r['avgchflow'][0, -1])))
        barrier()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device='cuda')
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        clocks = sampler.stop()
        ms = float(t[0])
        cm = float(ncell) * m * n_members
        nloc = max(er['stats']['members_local'], 1)
        line = _line('cell-months/s (PM+ABCD+MRTM)', 'cell-months/s', cm / (ms * 1e-3), ms / nloc, args, world_size,
                     {'workload': name, 'ncell': ncell, 'nmonths': m, 'members': n_members, 'nlcs': NLCS,
                      'runoff_spinup': RUNOFF_SPINUP, 'routing_spinup': ROUTING_SPINUP, 'dt_s': DT, 'scaling': 'strong',
                      'transport': 'float32 (forcing values are single precision; lossless)' if f32 else 'float64',
                      'parallelism': 'members over %d rank(s), %d on rank 0' % (world_size, nloc)},
                     {'e2e': {'value': cm / (ms * 1e-3), 'unit': 'cell-months/s', 'ms_per_step': ms / nloc,
                              'h2d_bytes_per_step': int(er['stats']['h2d_bytes'] // nloc),
                              'd2h_bytes_per_step': int(er['stats']['d2h_bytes'] // nloc),
                              'note': 'value IS the end-to-end number: every member comes from pinned host memory and its '
                                      'q / avgchflow / basin aggregates return to the host; ms_per_step = per member on a rank'},
                      'total_ms': ms, 'clocks': clocks, 'gpu_launches': 11 * nloc, 'timeline': er['stats'].get('timeline'),
                      'roofline': {'bound': 'hbm', 'kernel': 'pipeline (h2d | pm + abcd + mrtm | d2h)', 'achieved': None,
                                   'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': None, 'traffic': None,
                                   'h2d_gbs_per_rank': er['stats']['h2d_bytes'] / (ms * 1e-3) / 1e9,
                                   'd2h_gbs_per_rank': er['stats']['d2h_bytes'] / (ms * 1e-3) / 1e9}})
    else:
        raise SystemExit("unknown --workload %s" % name)
    if rank == 0:
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CPU baseline and `--impl reference`: the reference's OWN modules (oracle/_ref, staged unmodified by
# oracle/make_ref.sh), the numpy oracle port only if no staged tree travelled with the snapshot
# ------------------------------------------------------------------------------------------------
REF_SAMPLE = {'pm_cells': 16384, 'pm_months': 12, 'abcd_months': 36, 'mrtm_months': 2}


def _reference_modules():
    from oracle import ref_loader
    if ref_loader.staged_available():
        return ref_loader.load_staged(), 'reference'
    return None, 'port'


class ReferenceSample:
    """
    One bounded sample of the config-1 workload through the reference's stock code path:
      PM    run_pmpet (pet/penman_monteith.py:394) on the first `pm_cells` cells x 1 year, as Components calls it
            (single-threaded numpy);
      ABCD  abcd_execute (runoff/abcd.py:394) with jobs = -1 (joblib threads over 8 basin chunks, the reference's own
            parallelism) on ALL cells, `abcd_months` of spin-up + `abcd_months` of simulation;
      MRTM  downstream / upstream / upstream_genmatrix once (setup, not timed per step), then streamrouting
            (routing/mrtm.py:16) on ALL cells for `mrtm_months` months at dt = 3 h, carried state.
    cell-months/s of the whole pipeline = 1 / (t_pm + t_abcd (1 + spin-up / months) + t_mrtm (1 + spin-up / months))
    with t_* the measured seconds per cell-month-pass of each stage (every stage is linear in cell-months).
    """

    def __init__(self, world, pm, ab, nmonths):
        self.ref, self.kind = _reference_modules()
        self.world, self.pm, self.ab, self.nmonths = world, pm, ab, nmonths
        n = world.ncell
        nc = min(REF_SAMPLE['pm_cells'], n)
        self.nc = nc
        d = {}
        for k, v in pm.items():
            if isinstance(v, np.ndarray) and v.shape[:1] == (n,):
                v = v[:nc]
                if v.ndim == 2 and v.shape[1] >= 12 and k.endswith('_load') and k != 'lct_load':
                    v = v[:, :12]
            d[k] = v
        self.pm_sample = d
        ms = REF_SAMPLE['abcd_months']
        self.ms = ms
        self.pet_s = np.abs(np.random.default_rng(0).normal(80.0, 30.0, (n, ms)))
        self.q = np.abs(np.random.default_rng(0).normal(50.0, 30.0, (n, REF_SAMPLE['mrtm_months'])))
        self.ndays = month_days_mod4(12, START_YR)
        self._td = tempfile.TemporaryDirectory()
        self.calib_file = os.path.join(self._td.name, 'pars.npy')
        np.save(self.calib_file, ab['pars'])
        t0 = time.perf_counter()
        if self.kind == 'reference':
            s = world.settings()
            dsid = self.ref.mrtm.downstream(world.coords, world.flow_dir, s)
            upid = self.ref.mrtm.upstream(world.coords, dsid, s)
            self.um = self.ref.mrtm.upstream_genmatrix(upid)
        else:
            from oracle import mrtm as omrtm
            dsid = omrtm.downstream(world.coords, world.flow_dir, world.nrow, world.ncol)
            self.um = omrtm.csr_rows(omrtm.upstream_fast(world.coords, dsid, world.nrow, world.ncol))
        self.setup_s = time.perf_counter() - t0
        self.threads = 8                      # abcd_parallel: n_chunks = 8 for jobs < 1 (runoff/abcd.py:368-371)

    def step(self):
        w, n = self.world, self.world.ncell
        t0 = time.perf_counter()
        if self.kind == 'reference':
            data = SimpleNamespace(**{k: (np.copy(v) if isinstance(v, np.ndarray) else v)
                                      for k, v in self.pm_sample.items()})
            self.ref.pm.run_pmpet(data, self.nc, NLCS, START_YR, START_YR, self.pm['water_idx'], self.pm['snow_idx'],
                                  self.pm['lc_years'])
        else:
            from oracle import pet as opet
            opet.pm_pet(self.pm_sample, self.nc, NLCS, START_YR, START_YR, self.pm['water_idx'], self.pm['snow_idx'],
                        self.pm['lc_years'])
        t1 = time.perf_counter()
        ms = self.ms
        if self.kind == 'reference':
            self.ref.abcd.abcd_execute(n_basins=w.n_basins, basin_ids=w.basin_ids, pet=self.pet_s,
                                       precip=self.ab['precip'][:, :ms], tmin=self.ab['tmin'][:, :ms],
                                       calib_file=self.calib_file, n_months=ms, spinup_steps=ms, jobs=-1)
        else:
            from oracle import abcd as oabcd
            oabcd.abcd_execute(w.n_basins, w.basin_ids, self.pet_s, self.ab['precip'][:, :ms], self.ab['tmin'][:, :ms],
                               self.ab['pars'], ms, ms)
        t2 = time.perf_counter()
        S, F = np.zeros(n), np.zeros(n)
        nmr = self.q.shape[1]
        for m in range(nmr):
            if self.kind == 'reference':
                S, _, F = self.ref.mrtm.streamrouting(w.flow_dist, S, F, w.velocity, self.q[:, m], w.area,
                                                      int(self.ndays[m]), DT, self.um)
            else:
                from oracle import mrtm as omrtm
                S, _, F = omrtm.streamrouting(w.flow_dist, S, F, w.velocity, self.q[:, m], w.area, int(self.ndays[m]),
                                              DT, self.um)
        t3 = time.perf_counter()
        t_pm = (t1 - t0) / (self.nc * 12)
        t_abcd = (t2 - t1) / (n * 2 * ms)
        t_mrtm = (t3 - t2) / (n * nmr)
        spin_ro, spin_rt = min(RUNOFF_SPINUP, self.nmonths), min(ROUTING_SPINUP, self.nmonths)
        per_cm = t_pm + t_abcd * (1 + spin_ro / self.nmonths) + t_mrtm * (1 + spin_rt / self.nmonths)
        return {'wall_s': t3 - t0, 'value': 1.0 / per_cm,
                'stage_cell_months_per_s': {'pm': 1.0 / t_pm, 'abcd_per_pass': 1.0 / t_abcd,
                                            'mrtm_per_pass': 1.0 / t_mrtm}}

    def describe(self):
        return ('per step: run_pmpet 1 year x %d cells (1 thread, stock); abcd_execute(jobs=-1: 8 joblib threads over '
                'basin chunks) %d+%d months x %d cells; streamrouting %d months x %d cells at dt = %d s (scipy CSR, '
                '1 thread); stage rates combined to the %d-month workload with both spin-ups (every stage is linear '
                'in cell-months) - a sampled extrapolation, not a full run'
                % (self.nc, self.ms, self.ms, self.world.ncell, self.q.shape[1], self.world.ncell, DT, self.nmonths))


def cpu_baseline(world, pm, ab, nmonths):
    """One ReferenceSample step on the host cores (rank 0, N = 1), reported beside the GPU numbers."""
    rs = ReferenceSample(world, pm, ab, nmonths)
    r = rs.step()
    return {'value': r['value'], 'unit': 'cell-months/s', 'cores': rs.threads, 'kind': rs.kind,
            'sample': rs.describe(), 'stage_cell_months_per_s': r['stage_cell_months_per_s'],
            'sample_wall_s': r['wall_s'], 'host_cpus': os.cpu_count()}


def cpu_calib_baseline(world, pm, ab, n_eval=20):
    """objective_kge of the oracle port (one ABCD.emulate per candidate, calibrate_abcd.py:134-213) on the basin of
    median size, 1 thread; param-sets/s extrapolated to the mean basin size (cost is linear in cells)."""
    from oracle import calibrate as ocal
    counts = np.bincount(np.asarray(world.basin_ids).astype(int), minlength=world.n_basins + 1)[1:]
    b = int(np.argsort(counts)[len(counts) // 2]) + 1
    idx = np.nonzero(np.asarray(world.basin_ids) == b)[0]
    rng = np.random.default_rng(7)
    m = int(min(120, ab['precip'].shape[1]))
    pet = np.abs(rng.normal(80, 30, (len(idx), m)))
    precip, tmin = ab['precip'][idx, :m], ab['tmin'][idx, :m]
    obs = np.abs(rng.normal(1, 0.1, m))
    t0 = time.perf_counter()
    for k in range(n_eval):
        ocal.objective_kge(ab['pars'][b - 1] * (1 - 0.01 * k), pet, precip, tmin, m, m, 'km3_per_mth', world.area[idx], obs)
    dt_eval = (time.perf_counter() - t0) / n_eval
    per_cell_step = dt_eval / (len(idx) * 2 * m)
    mean_cells = float(counts.mean())
    return {'value': 1.0 / (per_cell_step * mean_cells * (NMONTHS + RUNOFF_SPINUP)), 'unit': 'param-sets/s', 'cores': 1,
            'kind': 'port', 'sample': '%d objective_kge evaluations on basin %d (%d cells), %d+%d months; extrapolated '
            'linearly to the mean basin (%.0f cells) and %d+%d months' % (n_eval, b, len(idx), m, m, mean_cells,
                                                                         RUNOFF_SPINUP, NMONTHS)}


def run_reference(args, rank, world_size):
    """
    --impl reference: the UNMODIFIED reference modules (oracle/_ref) on the host cores, rank 0 only.  A step is one
    bounded sample of the workload (ReferenceSample); `ms_per_step` is the wall time of that sample step, `value` the
    cell-months/s the stage rates of the step combine to for the full configuration.
    """
    if rank != 0:
        return
    world, pm, ab, end_yr = build_inputs(member_seed=1, ncell=args.ncell, nmonths=max(12, min(args.nmonths, 48)))
    rs = ReferenceSample(world, pm, ab, args.nmonths)
    res = []
    t_all = time.perf_counter()
    for k in range(args.warmup + args.steps):
        res.append(rs.step())
        if time.perf_counter() - t_all > 270 and k + 1 >= args.warmup + 1:
            break
    timed = res[min(args.warmup, len(res) - 1):]
    v = float(np.mean([x['value'] for x in timed]))
    wall = float(np.mean([x['wall_s'] for x in timed]))
    cb = {'value': v, 'unit': 'cell-months/s', 'cores': rs.threads, 'kind': rs.kind, 'sample': rs.describe(),
          'stage_cell_months_per_s': {k: float(np.mean([x['stage_cell_months_per_s'][k] for x in timed]))
                                      for k in timed[0]['stage_cell_months_per_s']},
          'host_cpus': os.cpu_count(), 'setup_s': rs.setup_s}
    line = {'impl': 'reference', 'metric': 'cell-months/s (PM+ABCD+MRTM)', 'value': v, 'unit': 'cell-months/s',
            'n_gpus': world_size, 'steps': len(timed), 'warmup': args.warmup, 'ms_per_step': wall * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': bench_config(args.ncell, args.nmonths),
            'note': 'ms_per_step is the wall time of one bounded sample step; value is the sampled extrapolation '
                    'described in cpu_baseline.sample',
            'cpu_baseline': cb,
            'e2e': {'value': v, 'unit': 'cell-months/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def _bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs next to its GPU, so that its pinned host buffers are allocated on that NUMA node and
    the per-step H2D / D2H copies of the end-to-end leg do not cross the socket link (best effort)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * i + b for i, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = ideal & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--ncell', type=int, default=NCELL)
    ap.add_argument('--nmonths', type=int, default=NMONTHS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-calib', action='store_true')
    ap.add_argument('--cpu-budget', type=float, default=20.0)
    ap.add_argument('--workload', default=WORKLOAD,
                    choices=[WORKLOAD, 'hs_abcd_1140', 'mrtm_hourly_360', 'calib_235x64x200', 'ensemble_64x1032'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world_size)
        return
    if world_size > 1:
        # stdout carries ONE JSON line.  NCCL prints its version banner / INFO log on the process's stdout: file
        # descriptor 1 is pointed at stderr for the whole run (the log is kept, e.g. for NCCL_DEBUG=INFO rank checks)
        # and the JSON line goes out through a duplicate of the original stdout.
        if 'XANTHOS_NCCL_DEBUG' in os.environ:
            os.environ['NCCL_DEBUG'] = os.environ['XANTHOS_NCCL_DEBUG']
        sys.stdout.flush()
        real_stdout = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)
        sys.stdout = real_stdout
        _bind_to_gpu_numa_node(local_rank)
        import torch
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        if args.workload != WORKLOAD:
            run_workload(args, rank, world_size, local_rank)
        else:
            run_ours(args, rank, world_size, local_rank)
    finally:
        if world_size > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
