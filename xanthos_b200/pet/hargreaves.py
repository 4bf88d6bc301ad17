"""
Hargreaves PET on the B200 - drop-in for xanthos/pet/hargreaves.py (the step-wise v1 PET).

`calculate_pet(temp, dtr, x, y, dr, m)` keeps the reference's per-month signature (hargreaves.py:17-39):
temperature and daily temperature range [ncell] of ONE month, latitude x [ncell] (radians), solar
declination y, inverse relative Earth-Sun distance dr and days in the month m (scalars); returns PET
[ncell] in mm/month and, like the reference, zeroes the negative entries of `dtr` in place.
`Components.simulation` does not loop over months: `series_device` evaluates the whole series in one
launch (every month is independent).
"""

import ctypes

import numpy as np

from .. import _cuda as C


def series_device(temp, dtr, lat_rad, solar_dec, dr, mth_days):
    """temp, dtr: [ncell, nmonths] host arrays or Fields; solar_dec, dr, mth_days: [nmonths] -> PET Field."""
    t, d = C.as_field(temp), C.as_field(dtr)
    lat = C.dev_vector(lat_rad)
    m = t.nmonths
    sd, sdp = C.as_c(np.asarray(solar_dec).reshape(-1)[:m], np.float64)
    rr, rrp = C.as_c(np.asarray(dr).reshape(-1)[:m], np.float64)
    dd, ddp = C.as_c(np.asarray(mth_days).reshape(-1)[:m], np.int32)
    if len(sd) != m or len(rr) != m or len(dd) != m:
        raise C.ValidationException("hargreaves: solar_dec / dr / mth_days must have one entry per month")
    pet = C.Field.empty(t.ncell, m, t.ld)
    C.check(C.lib().xan_hargreaves_pet(C.ptr(t.t), C.ptr(d.t), C.ptr(lat), sdp, rrp, ddp, C.ptr(pet.t), t.ncell, m,
                                       t.ld, C.stream_ptr()))
    return pet


def calculate_pet(temp, dtr, x, y, dr, m):
    """One month, reference signature (hargreaves.py:17-39)."""
    temp = np.asarray(temp, dtype=np.float64).reshape(-1)
    n = temp.shape[0]
    dtr_arr = np.asarray(dtr)
    if isinstance(dtr, np.ndarray):
        dtr[np.where(dtr < 0)[0]] = 0.            # the reference mutates its argument (hargreaves.py:32)
    pet = series_device(temp.reshape(n, 1), np.asarray(dtr_arr, dtype=np.float64).reshape(n, 1), x, [float(y)],
                        [float(dr)], [int(m)])
    return pet.to_host()[:, 0].copy()
