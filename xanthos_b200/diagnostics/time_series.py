"""
Aggregated time series - counterpart of xanthos/diagnostics/time_series.py.

`Aggregation_Map(Map, runoff)` (time_series.py:126-138) is an O(ncell x ntime) Python loop in the reference;
here it is one kernel (xan_group_sum) that adds the cells of every group in ascending cell index, i.e. in the
reference's own accumulation order - bit-identical.  Plotting (matplotlib) is not part of this library:
`TimeSeriesPlot` computes the series the reference plots and writes them as .csv next to where the .png would go.
"""

import logging
import os

import numpy as np

from .. import _cuda as C


def group_plan(id_map):
    """Cells sorted (stable) by group id, ids <= 0 left out -> (order int32 [n_used], offsets int32 [NB + 1], NB)."""
    ids = np.asarray(id_map).astype(np.int64).reshape(-1)
    nb = int(ids.max()) if ids.size else 0
    if nb < 1:
        return np.zeros(0, dtype=np.int32), np.zeros(1, dtype=np.int32), 0
    used = np.nonzero(ids > 0)[0]
    order = used[np.argsort(ids[used], kind='stable')]
    counts = np.bincount(ids[used], minlength=nb + 1)[1:]
    offsets = np.concatenate(([0], np.cumsum(counts)))
    return order.astype(np.int32), offsets.astype(np.int32), nb


_plans = {}   # id map (bytes digest) -> (device order, device offsets, NB): maps are few (basin, country, region)


def _device_plan(id_map):
    import hashlib
    torch = C.torch_cuda()
    ids = np.ascontiguousarray(np.asarray(id_map).astype(np.int64).reshape(-1))
    key = (ids.shape[0], hashlib.blake2b(ids.tobytes(), digest_size=16).digest(), torch.cuda.current_device())
    hit = _plans.get(key)
    if hit is None:
        order, offsets, nb = group_plan(ids)
        hit = (torch.from_numpy(order).cuda(), torch.from_numpy(offsets).cuda(), nb)
        if len(_plans) >= 16:
            _plans.clear()
        _plans[key] = hit
    return hit


def group_sum_device(id_map, t, ntime=None):
    """t: cuda tensor [ntime, ld] (time-major) -> cuda tensor [NB, ntime]."""
    torch = C.torch_cuda()
    d_order, d_off, nb = _device_plan(id_map)
    if nb < 1:
        raise C.ValidationException("Aggregation_Map: the id map has no positive id")
    ntime = int(t.shape[0]) if ntime is None else ntime
    out = torch.empty((nb, ntime), dtype=torch.float64, device='cuda')
    C.check(C.lib().xan_group_sum(C.ptr(t), C.ptr(d_order), C.ptr(d_off), nb, ntime, int(t.shape[1]), C.ptr(out),
                                  C.stream_ptr()))
    return out


def Aggregation_Map(Map, runoff):
    """[ncell, ntime] values -> [max(Map), ntime] sums over the cells of every id > 0, NaN skipped."""
    f = C.as_field(runoff)
    return group_sum_device(Map, f.t).cpu().numpy()


def TimeSeriesPlot(settings, Q, Avg_ChFlow, ref):
    """The aggregated runoff / streamflow series of time_series.py:20-67 as .csv (no matplotlib in this library)."""
    if not settings.CreateTimeSeriesPlot:
        return
    scales = {1: ['Basin'], 2: ['Country'], 3: ['GCAMRegion']}.get(settings.TimeSeriesScale,
                                                                    ['Basin', 'Country', 'GCAMRegion'])
    for scalestr in scales:
        key = {'Basin': 'basin', 'Country': 'country', 'GCAMRegion': 'region'}[scalestr]
        ids = getattr(ref, key + '_ids', None)
        names = getattr(ref, key + '_names', None)
        if ids is None:
            raise C.ValidationException("TimeSeriesPlot: no {} id map is loaded (reference data switched off?)".format(scalestr))
        if names is None:
            names = []
        folder = os.path.join(settings.OutputFolder, 'TimeSeriesPlot', scalestr)
        os.makedirs(folder, exist_ok=True)
        for arr, what in ((Q, 'runoff'), (Avg_ChFlow, 'streamflow')):
            agg = Aggregation_Map(ids, arr)
            agg = np.insert(agg, 0, np.sum(agg, axis=0), axis=0)      # row 0 = global (time_series.py:95-96)
            labels = np.insert(np.asarray(names, dtype=object), 0, 'Global')[:agg.shape[0]]
            with open(os.path.join(folder, '{}_{}.csv'.format(scalestr, what)), 'w') as fh:
                for i in range(agg.shape[0]):
                    fh.write('{},{},{}\n'.format(i, labels[i] if i < len(labels) else '', ','.join(repr(float(v)) for v in agg[i])))
        logging.info("Scale: {}, series written to {} (plots are not produced by xanthos_b200)".format(scalestr, folder))
