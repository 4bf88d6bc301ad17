"""
Penman-Monteith PET on the B200 - drop-in for xanthos/pet/penman_monteith.py.

`run_pmpet(data, ncells, nlcs, start_yr, end_yr, water_idx, snow_idx, land_cover_years)` has the
reference's signature (penman_monteith.py:394-477).  `data` carries the DataLoader attributes of
data_load.py:92-135: tair_load, TMIN_load, rhs_load, wind_load, rsds_load, rlds_load [N, M],
lct_load [N, nlcs, n_lc_years], elev [N, 1], per-class vectors and [nlcs, 12] tables.
`data.tairprev_load` is not read: the kernel takes the previous cell's temperature directly
(data_load.py:128-129).
"""

import ctypes

import numpy as np

from .. import _cuda as C

_VEC = ('cL', 'beta', 'rslimit', 'Tminopen', 'Tminclose', 'VPDclose', 'VPDopen', 'RBLmin', 'RBLmax', 'rc', 'emiss')
_TAB = ('alpha', 'lai', 'laimin', 'laimax')


def land_cover_index(target_yr, land_cover_years):
    """Index of the land-cover slice used for a year (SetData, penman_monteith.py:32-43)."""
    lc = sorted(land_cover_years)
    if target_yr >= lc[-1]:
        return lc.index(lc[-1])
    return lc.index([x for x in lc if x - target_yr >= -4][0])


def stage_land_cover(lct_load, ld):
    """[N, nlcs, nyears] host array -> device tensor [nyears, nlcs, ld] (layout change only)."""
    torch = C.torch_cuda()
    if isinstance(lct_load, torch.Tensor) and lct_load.is_cuda and lct_load.dim() == 3 and lct_load.shape[2] == ld:
        return lct_load
    a = torch.from_numpy(np.ascontiguousarray(lct_load, dtype=np.float64)).cuda()
    n, nlcs, ny = a.shape
    out = torch.zeros((ny, nlcs, ld), dtype=torch.float64, device='cuda')
    out[:, :, :n] = a.permute(2, 1, 0)
    return out


def run_pmpet_device(data, ncells, nlcs, start_yr, end_yr, water_idx, snow_idx, land_cover_years, prev_idx=None):
    nlcs = int(nlcs)
    if nlcs < 7:
        raise IndexError("index 6 is out of bounds for axis 0 with size {}".format(nlcs))   # reference :377
    f = [C.as_field(getattr(data, k), nan_to_num=False) for k in
         ('tair_load', 'TMIN_load', 'rhs_load', 'wind_load', 'rsds_load', 'rlds_load')]
    tair = f[0]
    nyears = int(end_yr) - int(start_yr) + 1
    if tair.nmonths != nyears * 12:
        raise C.ValidationException("PM forcing has {} months, expected {}".format(tair.nmonths, nyears * 12))
    lct = stage_land_cover(data.lct_load, tair.ld)
    elev = C.dev_vector(data.elev)
    keep = []
    tabs = C.PmTables()
    tabs.nlcs, tabs.water_idx, tabs.snow_idx = nlcs, int(water_idx), int(snow_idx)
    for k in _VEC:
        a, p = C.as_c(np.asarray(getattr(data, k)).reshape(-1)[:nlcs], np.float64)
        keep.append(a)
        setattr(tabs, k, p)
    for k in _TAB:
        a, p = C.as_c(np.asarray(getattr(data, k))[:nlcs, :12], np.float64)
        keep.append(a)
        setattr(tabs, k, p)
    lc_idx, lcp = C.as_c(np.array([land_cover_index(y, land_cover_years) for y in range(int(start_yr), int(end_yr) + 1)]),
                         np.int32)
    pidx = None if prev_idx is None else C.dev_vector(prev_idx, np.int32)
    pet = C.Field.empty(tair.ncell, tair.nmonths, tair.ld)
    C.check(C.lib().xan_pm_pet(C.ptr(f[0].t), C.ptr(f[1].t), C.ptr(f[2].t), C.ptr(f[3].t), C.ptr(f[4].t),
                               C.ptr(f[5].t), C.ptr(lct), C.ptr(elev), C.ptr(pidx), ctypes.byref(tabs), lcp,
                               C.ptr(pet.t), tair.ncell, tair.nmonths, tair.ld, int(start_yr), C.stream_ptr()))
    return pet


def run_pmpet(data, ncells, nlcs, start_yr, end_yr, water_idx, snow_idx, land_cover_years):
    pet = run_pmpet_device(data, ncells, nlcs, start_yr, end_yr, water_idx, snow_idx, land_cover_years)
    return C.remember(pet.to_host(), pet)
