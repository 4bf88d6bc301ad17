"""
Day-count tables (oracle; test infrastructure only).

Three leap-year conventions coexist in the reference and all are reproduced:
  * mod-4 rule (2100 is leap):  xanthos/utils/general.py:15-50  -> MRTM `nday`
  * Gregorian rule (calendar):  xanthos/pet/penman_monteith.py:57-62,
    xanthos/pet/thornthwaite.py:113, xanthos/pet/hargreaves_samani.py:18-28
"""

import calendar

import numpy as np

MONTHDAYS = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)
LEAP_MONTHDAYS = (31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)


def set_month_arrays(n_months, start_year, end_year):
    """[year, month index, days] per month with the `year % 4 == 0` rule (general.py:15-50)."""
    out = np.zeros((n_months, 3), dtype=int)
    k = 0
    for y in range(start_year, end_year + 1):
        days = LEAP_MONTHDAYS if y % 4 == 0 else MONTHDAYS
        for j in range(12):
            out[k] = (y, j, days[j])
            k += 1
    return out


def gregorian_days(start_year, end_year):
    """Days per month with calendar.monthrange (hargreaves_samani.py:18-28)."""
    return np.array([calendar.monthrange(y, m)[1]
                     for y in range(start_year, end_year + 1) for m in range(1, 13)], dtype=float)
