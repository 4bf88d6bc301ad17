"""
Multi-member runs of the hot path (BASELINE.json config 5: "64-member ESM x SSP scenario ensemble ... sharded by member").

The reference has no ensemble driver: a user loops over configuration files, one `Xanthos(ini).execute()` per scenario
(xanthos/model.py, components.py:298-385), and every scenario re-reads the static data.  Here the static part (land
cover, PM tables, basin plan, routing plan, ABCD parameters) is staged once and the members stream through
PET -> ABCD -> MRTM with their copies overlapped:

    h2d stream     : forcing of the next members  (8 fields each, pinned host -> HBM; transposed to month-major on the
                     compute stream)
    compute stream : transposes, PM -> ABCD of members k and k+1, then ONE routing launch for both
                     (`route_device_batch`: the thread blocks of the two members share the SMs), basin aggregates,
                     output transposes.  `XANTHOS_ENSEMBLE_FRONT=1` moves PM -> ABCD of the NEXT pair to a stream of
                     its own, beside the routing launch: measured slower (75.7 against 72.0 ms per pair - every block
                     that shares an SM with the latency-bound routing chain slows it by more than it saves), so it is off.
    d2h stream     : requested outputs of the members before  (HBM -> pinned host, cell-major like the reference's arrays)

The forcing of at most `prefetch_depth` (4 = two pairs) members is on its way ahead of the group being computed, into a
fixed ring of device staging buffers (no allocation inside the pipeline).  Only the variables named in `output_vars` are copied back (the reference
keeps PET, AET, Q, Sav, ChStorage and Avg_ChFlow of a scenario in host memory but writes `output_vars` only,
data_writer/out_writer.py:60-110).  With torch.distributed initialised the members are dealt in contiguous blocks to the ranks
(no collective in the data path) and the basin aggregates [n_members, 2, nmonths, n_basins] are gathered at the end.
"""

from types import SimpleNamespace

import numpy as np

from . import _cuda as C
from . import sharding
from .pet import penman_monteith as pm_mod
from .runoff import abcd as abcd_mod
from .routing import mrtm as mrtm_mod

PM_FORCING = ('tair_load', 'TMIN_load', 'rhs_load', 'wind_load', 'rsds_load', 'rlds_load')
FORCING = PM_FORCING + ('precip', 'tmin')
OUTPUTS = ('pet', 'aet', 'q', 'soilmoisture', 'chstorage', 'avgchflow')


def lossless_float32(member, pin=True):
    """
    Returns the member with every forcing array whose values are all exactly representable in single precision replaced
    by a float32 copy (pinned when a device is present); the others are kept.  Climate-model forcing is single precision
    on disk (NetCDF float variables) and only widened to float64 by the loader (data_load.py:288-340), so the copy is
    lossless: float -> double on the device is exact and every result stays bit-identical, while the host -> device
    traffic - the limiter of multi-GPU ensemble runs - halves.  NaN (missing data) counts as representable.
    Call it once per member where the data are loaded; it costs a pass over the arrays on the host.
    """
    out = dict(member)
    for k in FORCING:
        a = np.asarray(member[k])
        if a.dtype != np.float64:
            continue
        a32 = a.astype(np.float32)
        back = a32.astype(np.float64)
        if np.array_equal(back, a, equal_nan=True):
            if pin and C.device_available():
                t = C.torch_cuda().from_numpy(a32).pin_memory()
                a32 = t.numpy()
            out[k] = a32
    return out


class EnsembleStatics:
    """Everything that does not change from member to member, staged on the device once."""

    def __init__(self, ncell, start_yr, end_yr, pm_tables, lct_load, elev, water_idx, snow_idx, lc_years, nlcs,
                 n_basins, basin_ids, abcd_pars, area, flow_dist, velocity, upstream_matrix, ndays, dt=3 * 3600,
                 runoff_spinup=None, routing_spinup=None):
        torch = C.torch_cuda()
        self.ncell, self.start_yr, self.end_yr = int(ncell), int(start_yr), int(end_yr)
        self.nmonths = (self.end_yr - self.start_yr + 1) * 12
        self.ld = C.padded_ld(self.ncell)
        self.tables = dict(pm_tables)
        self.water_idx, self.snow_idx, self.lc_years, self.nlcs = int(water_idx), int(snow_idx), list(lc_years), int(nlcs)
        self.d_lct = pm_mod.stage_land_cover(lct_load, self.ld)
        self.d_elev = C.dev_vector(elev)
        self.n_basins = int(n_basins)
        self.basin_ids = np.asarray(basin_ids)
        rows = abcd_mod._basin_rows(self.n_basins, self.basin_ids, np.asarray(abcd_pars).shape[0])
        self.plan = abcd_mod.basin_plan(rows, self.n_basins)
        self.d_pars = torch.from_numpy(np.ascontiguousarray(abcd_pars, dtype=np.float64)).cuda()
        self.d_area, self.d_L, self.d_V = C.dev_vector(area), C.dev_vector(flow_dist), C.dev_vector(velocity)
        self.d_area_km3 = self.d_area * 1e-6
        self.um = upstream_matrix
        self.ndays = np.asarray(ndays, dtype=np.int32)[:self.nmonths]
        self.dt = float(dt)
        self.runoff_spinup = self.nmonths if runoff_spinup is None else int(runoff_spinup)
        self.routing_spinup = self.nmonths if routing_spinup is None else int(routing_spinup)

    def data_ns(self, fields):
        d = dict(self.tables)
        d.update({k: fields[k] for k in PM_FORCING})
        d['lct_load'] = self.d_lct
        d['elev'] = self.d_elev
        return SimpleNamespace(**d)


class EnsembleRunner:
    def __init__(self, statics, output_vars=('q', 'avgchflow'), aggregates=True, prefetch_depth=4, group=2):
        torch = C.torch_cuda()
        self.prefetch_depth = max(1, int(prefetch_depth))   # members whose forcing may be on its way ahead of the compute
        self.group = max(1, int(group))                     # members routed by one launch
        bad = [v for v in output_vars if v not in OUTPUTS]
        if bad:
            raise C.ValidationException("unknown output variable(s) {}; choose from {}".format(bad, OUTPUTS))
        self.s, self.output_vars, self.aggregates = statics, tuple(output_vars), bool(aggregates)
        # One pair of copy streams per process and device, not per runner: torch's caching allocator keeps freed blocks
        # per STREAM, so every new stream starts with an empty pool - a runner per run_ensemble call with fresh streams
        # left 6 - 9 GB of staging blocks cached under dead streams per call, and after a dozen calls the allocator had
        # to cudaFree them (a device-wide synchronisation) in the middle of the uploads: 250 - 450 ms per stalled upload.
        self.h2d, self.d2h = C.copy_stream('ensemble_h2d'), C.copy_stream('ensemble_d2h')
        import os as _os
        self.front = C.copy_stream('ensemble_front') if _os.environ.get('XANTHOS_ENSEMBLE_FRONT', '0') == '1' else None
        self.h2d_bytes = self.d2h_bytes = 0
        self._slots = None          # ring of device staging buffers for the uploads, depth + group slots
        self._torch = torch
        import os
        self.timeline = [] if os.environ.get('XANTHOS_ENSEMBLE_TIMELINE') else None   # per member: stage event pairs

    # ---- the three pipeline stages (everything is enqueued, nothing waits) ------------------------------------------
    def _slot_buffer(self, slot, k):
        """Device staging buffer of forcing field k in ring slot `slot` (made on first use, then reused for the whole
        run: a `cudaMalloc` per upload now and then stalls for hundreds of ms while a routing launch is running)."""
        torch = self._torch
        if self._slots is None or len(self._slots) <= slot:
            self._slots = (self._slots or []) + [dict() for _ in range(slot + 1 - len(self._slots or []))]
        b = self._slots[slot].get(k)
        if b is None:
            b = self._slots[slot][k] = torch.empty(self.s.ncell * self.s.nmonths, dtype=torch.float64, device='cuda')
        return b

    def _upload(self, member, slot):
        torch = self._torch
        if callable(member):
            member = member()
        missing = [k for k in FORCING if k not in member]
        if missing:
            raise C.ValidationException("ensemble member lacks {}".format(missing))
        staged = {}
        bufs = {k: self._slot_buffer(slot, k) for k in FORCING}      # allocated on the compute stream (its pool is reused)
        # Only copy-engine work goes on the h2d stream.  The transposes to month-major run on the compute stream in front
        # of the member's kernels: a kernel on the h2d stream would wait for SM resources behind the routing kernel of
        # the previous member (one cooperative block per SM, the whole register file) and hold up the next copy.
        with torch.cuda.stream(self.h2d):
            e0 = self._mark(self.h2d)
            for k in FORCING:
                a = member[k]
                if tuple(a.shape) != (self.s.ncell, self.s.nmonths):
                    raise C.ValidationException("member field {} has shape {}, expected {}".format(
                        k, tuple(a.shape), (self.s.ncell, self.s.nmonths)))
                if not isinstance(a, torch.Tensor):
                    a = np.asarray(a)
                    if a.dtype not in (np.float64, np.float32) or not a.flags['C_CONTIGUOUS']:
                        a = np.ascontiguousarray(a, dtype=np.float64)
                    a = torch.from_numpy(a)
                if a.dtype not in (torch.float64, torch.float32):
                    a = a.to(torch.float64)
                # float32 arrays cross the link as float32 (see lossless_float32).  The destination is the ring slot's
                # buffer: the slot is free, `run` has waited for the compute stage of its previous occupant
                buf = bufs[k]
                t = (buf.view(torch.float32)[:a.numel()] if a.dtype == torch.float32 else buf).view(a.shape)
                t.copy_(a, non_blocking=True)
                staged[k] = t
                self.h2d_bytes += t.numel() * t.element_size()
            ev = torch.cuda.Event(enable_timing=self.timeline is not None)
            ev.record(self.h2d)
            if self.timeline is not None:
                self.timeline.append({'h2d': (e0, ev)})
        return staged, ev

    def _to_fields(self, staged):
        """cell-major staging tensors -> month-major Fields (compute stream)."""
        fields = {}
        for k, t in staged.items():
            f = C.Field.empty(self.s.ncell, self.s.nmonths, self.s.ld)
            fn = C.lib().xan_to_month_major_f32 if t.dtype == self._torch.float32 else C.lib().xan_to_month_major
            C.check(fn(C.ptr(t), C.ptr(f.t), self.s.ncell, self.s.nmonths, f.ld, 0, C.stream_ptr()))
            fields[k] = f
        return fields

    def _mark(self, stream):
        if self.timeline is None:
            return None
        e = self._torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    def _front(self, uploads):
        """Front stage, per member: wait for its upload, transposes, PM -> ABCD.  `uploads`: [(staged tensors, upload event)].
        Runs on the front stream (beside the routing launch of the group before); returns (PET Fields, ABCD results, event)."""
        s, torch = self.s, self._torch
        compute = torch.cuda.current_stream()
        stream = self.front if self.front is not None else compute
        want = tuple(k for k, v in (('aet', 'aet'), ('q', 'q'), ('sav', 'soilmoisture')) if v in self.output_vars or k == 'q')
        pets, ress = [], []
        with torch.cuda.stream(stream):
            while uploads:
                staged, ev = uploads.pop(0)
                stream.wait_event(ev)     # the first member's PET and runoff run under the second member's upload
                fields = self._to_fields(staged)
                del staged
                pet = pm_mod.run_pmpet_device(s.data_ns(fields), s.ncell, s.nlcs, s.start_yr, s.end_yr, s.water_idx,
                                              s.snow_idx, s.lc_years)
                res = abcd_mod.run_device(s.plan, s.d_pars, pet, fields['precip'], fields['tmin'], s.nmonths,
                                          s.runoff_spinup, want=want)
                if stream is not compute:     # allocated on the front stream, read by the routing / output stage
                    for f in [pet] + list(res.values()):
                        f.t.record_stream(compute)
                pets.append(pet)
                ress.append(res)
            ev = torch.cuda.Event()
            ev.record(stream)
        return pets, ress, ev

    def _back(self, pets, ress):
        """One routing call for the group and the basin aggregates per member (compute stream)."""
        s = self.s
        want_chs = 'chstorage' in self.output_vars
        if len(ress) == 1:
            routed = [mrtm_mod.route_device(s.um, ress[0]['q'], s.d_L, s.d_V, s.d_area, s.ndays, s.dt, s.routing_spinup,
                                            want_chs=want_chs)]
        else:
            routed = mrtm_mod.route_device_batch(s.um, [r['q'] for r in ress], s.d_L, s.d_V, s.d_area, s.ndays, s.dt,
                                                 s.routing_spinup, want_chs=want_chs)
        results = []
        for pet, res, (chs, avg, _) in zip(pets, ress, routed):
            out = {'pet': pet, 'aet': res.get('aet'), 'q': res['q'], 'soilmoisture': res.get('sav'), 'chstorage': chs,
                   'avgchflow': avg}
            agg = None
            if self.aggregates:   # basin runoff in km3 / month and basin sum of the mean streamflow, [2, nmonths, n_basins]
                agg = self._torch.empty((2, s.nmonths, s.n_basins), dtype=self._torch.float64, device='cuda')
                C.check(C.lib().xan_basin_sum(s.plan._plan, C.ptr(res['q'].t), C.ptr(s.d_area_km3), s.nmonths,
                                              res['q'].ld, C.ptr(agg[0]), C.stream_ptr()))
                C.check(C.lib().xan_basin_sum(s.plan._plan, C.ptr(avg.t), None, s.nmonths, avg.ld, C.ptr(agg[1]),
                                              C.stream_ptr()))
            results.append((out, agg))
        return results

    def _download(self, out, agg):
        torch = self._torch
        compute = torch.cuda.current_stream()
        staged = {v: out[v].to_device_cell_major() for v in self.output_vars}      # transposes run on the compute stream
        ev = torch.cuda.Event()
        ev.record(compute)
        self.d2h.wait_event(ev)
        host = {}
        with torch.cuda.stream(self.d2h):
            for v, dev in staged.items():
                dev.record_stream(self.d2h)
                h = C.host_pool.acquire((self.s.ncell, self.s.nmonths))
                h.copy_(dev, non_blocking=True)
                host[v] = h
                self.d2h_bytes += h.numel() * 8
            if agg is not None:
                agg.record_stream(self.d2h)
                h = torch.empty(tuple(agg.shape), dtype=torch.float64, pin_memory=True)
                h.copy_(agg, non_blocking=True)
                host['basin_aggregates'] = h
                self.d2h_bytes += h.numel() * 8
            done = torch.cuda.Event(enable_timing=self.timeline is not None)
            done.record(self.d2h)
        return host, done

    # ---- driver ----------------------------------------------------------------------------------------------------------
    def run(self, members):
        """Generator: yields (index, {variable: host ndarray [ncell, nmonths], 'basin_aggregates': [2, nmonths, n_basins]})
        in member order; the next members are being uploaded and a group of `group` members computed while the outputs of
        the group before are handed out."""
        torch = self._torch
        members = list(members)
        if not members:
            return
        compute = torch.cuda.current_stream()
        n, depth, g = len(members), self.prefetch_depth, self.group
        # pinned output buffers in flight: the group being computed, the group before it on its way to the host, and the
        # results the caller still holds from the last hand-out - allocated before the pipeline starts, not inside it
        C.host_pool.reserve((self.s.ncell, self.s.nmonths), min(3 * g, n + g) * len(self.output_vars))
        uploaded, next_up = {}, 0
        pending = []              # (index, host tensors, done event) of the members whose outputs are in flight
        computed = {}             # member -> completion event of its group's compute stage, to bound the uploads' run-ahead

        def pump(k):              # uploads of the members up to k + depth; member j waits for the compute of member j - depth - g
            nonlocal next_up
            while next_up < n and next_up <= k + depth:
                old = next_up - depth - g
                if old >= 0:
                    computed[old].synchronize()
                uploaded[next_up] = self._upload(members[next_up], next_up % (depth + g))
                next_up += 1
        fronts = {}

        def start_front(k0):      # front stage of the group starting at member k0
            ks = list(range(k0, min(n, k0 + g)))
            if ks[-1] not in uploaded:
                pump(ks[-1] - depth)
            fronts[k0] = self._front([uploaded.pop(k) for k in ks])
        pump(g - 1)
        start_front(0)
        for k0 in range(0, n, g):
            ks = list(range(k0, min(n, k0 + g)))
            pets, ress, fev = fronts.pop(k0)
            c0 = self._mark(compute)
            compute.wait_event(fev)
            results = self._back(pets, ress)
            del pets, ress
            if k0 + g < n:        # enqueued behind this group's routing launch, runs beside it on the device
                start_front(k0 + g)
            downloads = [self._download(out, agg) for out, agg in results]
            del results
            cev = torch.cuda.Event(enable_timing=self.timeline is not None)
            cev.record(compute)
            for k, (host, done) in zip(ks, downloads):
                computed[k] = cev
                if self.timeline is not None:
                    self.timeline[k]['compute'] = (c0, cev)
                    self.timeline[k]['done'] = done
            pump(ks[-1])
            for p in pending:
                yield self._finish(p)
            pending = [(k, host, done) for k, (host, done) in zip(ks, downloads)]
        for p in pending:
            yield self._finish(p)

    @staticmethod
    def _finish(p):
        k, host, done = p
        done.synchronize()
        res = {}
        for v, h in host.items():
            res[v] = C.host_pool.as_array(h) if v != 'basin_aggregates' else h.numpy()
        return k, res


def run_ensemble(statics, members, output_vars=('q', 'avgchflow'), aggregates=True, on_result=None):
    """
    Run every member (this rank's share when torch.distributed is initialised: a contiguous block per rank,
    sharding.partition_members).
    Returns {member index: result dict}; with `on_result(index, result)` given the results are handed to it instead of
    being kept (a 64-member ensemble of 1,032 months is 2 x 36 GB of output).  On every rank the returned dict also has
    the key 'basin_aggregates': [n_members, 2, nmonths, n_basins] of ALL members (all-gather over NCCL).
    """
    torch = C.torch_cuda()
    import torch.distributed as dist
    members = list(members)
    dist_on = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist_on else (0, 1)
    mine = [int(i) for i in sharding.partition_members(len(members), world)[rank]]
    runner = EnsembleRunner(statics, output_vars, aggregates)
    results = {}
    agg_local = torch.zeros((len(mine), 2, statics.nmonths, statics.n_basins), dtype=torch.float64)
    for j, res in runner.run([members[i] for i in mine]):
        if aggregates:
            agg_local[j] = torch.from_numpy(res['basin_aggregates'])
        if on_result is not None:
            on_result(mine[j], res)
        else:
            results[mine[j]] = res
    if aggregates:      # the only collective: every rank gets the basin aggregates of all members
        results['basin_aggregates'] = sharding.gather_ragged_rows(mine, agg_local.cuda(), len(members)).cpu().numpy()
    results['stats'] = {'h2d_bytes': runner.h2d_bytes, 'd2h_bytes': runner.d2h_bytes, 'members_local': len(mine)}
    if runner.timeline:      # ms relative to the first upload: (h2d start, end), (compute start, end), outputs on the host
        torch.cuda.synchronize()
        t0 = runner.timeline[0]['h2d'][0]
        results['stats']['timeline'] = [
            {'h2d': [round(t0.elapsed_time(e), 1) for e in m['h2d']],
             'compute': [round(t0.elapsed_time(e), 1) for e in m['compute']],
             'done': round(t0.elapsed_time(m['done']), 1)} for m in runner.timeline if 'compute' in m]
    return results
