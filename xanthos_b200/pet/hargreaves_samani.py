"""
Hargreaves-Samani PET on the B200 - drop-in for xanthos/pet/hargreaves_samani.py.

`execute(config, data, out_file=None)` has the reference's signature (hargreaves_samani.py:91-119):
`config` supplies ncell / nmonths / StartYear / EndYear, `data` supplies hs_tas / hs_tmax / hs_tmin
[ncell, nmonths] and coords (latitude in column 2).  Returns PET [ncell, nmonths] in mm/month.
"""

import numpy as np

from .. import _cuda as C


def execute_device(tas, tmax, tmin, lat_deg, start_year):
    """Fields (or host arrays) in, PET Field out; nothing leaves the device."""
    t, hi, lo = C.as_field(tas), C.as_field(tmax), C.as_field(tmin)
    lat = C.dev_vector(lat_deg)
    pet = C.Field.empty(t.ncell, t.nmonths, t.ld)
    C.check(C.lib().xan_hs_pet(C.ptr(t.t), C.ptr(hi.t), C.ptr(lo.t), C.ptr(lat), C.ptr(pet.t), t.ncell, t.nmonths,
                               t.ld, int(start_year), C.stream_ptr()))
    return pet


def execute(config, data, out_file=None):
    pet = execute_device(data.hs_tas, data.hs_tmax, data.hs_tmin, data.coords[:, 2], config.StartYear)
    out = C.remember(pet.to_host(), pet)
    if out_file is not None:
        np.save(out_file, out)
    return out
