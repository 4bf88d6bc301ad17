"""
Thornthwaite PET on the B200 - drop-in for xanthos/pet/thornthwaite.py.

`execute(tas, lat_radians, start_yr, end_yr)` as in the reference (thornthwaite.py:47-130);
`calc_daylight_hours` is kept for API parity (thornthwaite.py:18-44).  Unlike the reference the
input `tas` is NOT modified in place (the reference zeroes NaN / negative entries, :82; the kernel
applies the same rule on the fly).
"""

import numpy as np

from .. import _cuda as C


def execute_device(tas, lat_radians, start_yr):
    t = C.as_field(tas)
    lat = C.dev_vector(lat_radians)
    pet = C.Field.empty(t.ncell, t.nmonths, t.ld)
    C.check(C.lib().xan_thornthwaite_pet(C.ptr(t.t), C.ptr(lat), C.ptr(pet.t), t.ncell, t.nmonths, t.ld,
                                         int(start_yr), C.stream_ptr()))
    return pet


def execute(tas, lat_radians, start_yr, end_yr):
    nmonths = (int(end_yr) - int(start_yr) + 1) * 12
    if not isinstance(tas, C.Field) and np.asarray(tas).shape[1] != nmonths:
        raise C.ValidationException("tas has {} months, expected {}".format(np.asarray(tas).shape[1], nmonths))
    pet = execute_device(tas, lat_radians, start_yr)
    return C.remember(pet.to_host(), pet)


def calc_daylight_hours(mth_days, lat_radians):
    """
    Monthly mean day length in hours [n_lat, 12] (thornthwaite.py:18-44).  `mth_days` must be the
    month lengths of a 365- or a 366-day year (the only two the reference ever passes, :109-115).
    """
    torch = C.torch_cuda()
    days = [int(d) for d in mth_days]
    regular = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    leap = [31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    if days not in (regular, leap):
        raise C.ValidationException("calc_daylight_hours supports 365- and 366-day years only")
    lat = C.dev_vector(lat_radians)
    n = lat.shape[0]
    out = torch.empty((24, n), dtype=torch.float64, device='cuda')
    C.check(C.lib().xan_thornthwaite_daylight(C.ptr(lat), C.ptr(out), n, C.stream_ptr()))
    rows = out[12:] if days == leap else out[:12]
    return rows.t().contiguous().cpu().numpy()
