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
