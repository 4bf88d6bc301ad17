"""
General helper functions kept from xanthos/utils/general.py.

`set_month_arrays` defines the MRTM sub-step count per month and therefore keeps the reference's
`year % 4 == 0` leap rule (general.py:15-50); 2100 counts as a leap year here while the PET modules
use the Gregorian calendar.
"""

import numpy as np

_M1 = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)
_M2 = (31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)


def set_month_arrays(n_months, start_year, end_year):
    """[[year, month index, days in month], ...] as an int array [n_months, 3] (general.py:15-50)."""
    out = np.zeros((n_months, 3), dtype=int)
    k = 0
    for y in range(start_year, end_year + 1):
        days = _M2 if y % 4 == 0 else _M1
        for j in range(12):
            if k >= n_months:
                return out
            out[k] = (y, j, days[j])
            k += 1
    return out


def calc_sinusoidal_factor(yr_imth_ndays, startmonth=1):
    """
    Monthly means of the solar declination (radians) and of the inverse relative Earth-Sun distance
    (general.py:53-90): daily values 0.409 sin(2 pi j / N - 1.39 + ph) and 1 + 0.033 cos(2 pi j / N + ph)
    averaged over the days of the month, N = 365 or 366 by the `year % 4` rule.  Host-side: 2 x nmonths
    scalars that parameterise the Hargreaves kernel.  Returns (solar_dec, dr).
    """
    ymd = np.asarray(yr_imth_ndays)
    n = ymd.shape[0]
    solar_dec, dr = np.zeros(n), np.zeros(n)
    ph = (startmonth - 1.) / 12. * 2. * np.pi
    cache = {}
    for i in range(n):
        leap = int(ymd[i, 0]) % 4 == 0
        if leap not in cache:
            mdays = _M2 if leap else _M1
            ends = np.cumsum(mdays)
            j = np.arange(1, ends[-1] + 1)
            cache[leap] = (ends - np.asarray(mdays), ends,
                           0.409 * np.sin(2 * np.pi * j / max(j) - 1.39 + ph),
                           1. + 0.033 * np.cos(2 * np.pi * j / max(j) + ph))
        beg, end, lam, d = cache[leap]
        mth = int(ymd[i, 1])
        solar_dec[i] = np.mean(lam[beg[mth]:end[mth]])
        dr[i] = np.mean(d[beg[mth]:end[mth]])
    return solar_dec, dr
