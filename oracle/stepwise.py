"""
Oracle of the step-wise (legacy v1) path: Hargreaves PET and GWAM runoff.  TEST INFRASTRUCTURE ONLY.

Restates, in plain numpy:
  * `calc_sinusoidal_factor`   xanthos/utils/general.py:53-90
  * `hargreaves_pet`           xanthos/pet/hargreaves.py:17-73 (calculate_pet, calc_insolation,
                               calc_daylight_hours and the clipping `acos`)
  * `gwam_runoffgen`           xanthos/runoff/gwam.py:18-88 (runoffgen)
  * `stepwise_run`             the month loops of Components.simulation (components.py:329-366) as driven
                               by ConfigRunner.run (configurations.py:106-123): a spin-up pass over the
                               first `runoff_spinup` months that only carries the soil moisture over,
                               then the simulation from month 0.
Pinned bitwise to the reference by oracle/validate_against_reference.py and tests/golden/case_c.npz.
"""

import numpy as np


def calc_sinusoidal_factor(yr_imth_ndays, startmonth=1):
    """Monthly means of solar declination and inverse relative Earth-Sun distance (general.py:53-90)."""
    n = yr_imth_ndays.shape[0]
    solar_dec = np.zeros(n)
    dr = np.zeros(n)
    first = {True: np.array([1, 32, 61, 92, 122, 153, 183, 214, 245, 275, 306, 336]),     # :66
             False: np.array([1, 32, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335])}    # :74
    last = {True: np.array([31, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335, 366]),    # :67
            False: np.array([31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334, 365])}   # :75
    ph = (startmonth - 1.) / 12. * 2. * np.pi                                              # :78
    for i in range(n):
        leap = np.mod(yr_imth_ndays[i, 0], 4) == 0                                         # :64 (mod-4 rule)
        j = np.arange(1, 367 if leap else 366)
        lam = 0.409 * np.sin(2 * np.pi * j / max(j) - 1.39 + ph)                           # :79
        d = 1. + 0.033 * np.cos(2 * np.pi * j / max(j) + ph)                               # :80
        mth = yr_imth_ndays[i, 1]
        solar_dec[i] = np.mean(lam[first[leap][mth] - 1:last[leap][mth]])                  # :84
        dr[i] = np.mean(d[first[leap][mth] - 1:last[leap][mth]])                           # :85
    return solar_dec, dr


def clipped_acos(x):
    """hargreaves.py:53-73: arccos with arguments beyond [-1, 1] mapped to pi / 0; NaN -> 0."""
    y = np.zeros_like(x)
    inside = (x <= 1) & (x >= -1)
    y[inside] = np.arccos(x[inside])
    y[x < -1] = np.arccos(-1)
    y[x > 1] = np.arccos(1)
    return y


def hargreaves_pet(temp, dtr, lat_rad, solar_dec, dr, mth_days):
    """One month of Hargreaves PET [ncell], mm/month (hargreaves.py:17-39). `dtr` is not modified."""
    ws = clipped_acos(-np.tan(lat_rad) * np.tan(solar_dec))                                # :48-50
    ra = 15.392 * dr * (ws * np.sin(lat_rad) * np.sin(solar_dec)
                        + np.cos(lat_rad) * np.cos(solar_dec) * np.sin(ws))                # :42-44
    d = np.array(dtr, dtype=float)
    d[d < 0] = 0.                                                                          # :32
    evap = mth_days * 0.0023 * ra * (temp + 17.8) * np.sqrt(d)                             # :35
    return np.maximum(evap, np.zeros_like(evap))                                           # :36


def gwam_runoffgen(pet, p, sm, chstor, indexing=999):
    """One month of GWAM (gwam.py:18-88) -> (aet, q, sav), each [ncell]."""
    n = pet.shape[0]
    b = chstor + p - pet                                                                   # :42
    sav, q, aet = np.zeros(n), np.zeros(n), np.zeros(n)
    lake = sm == indexing                                                                  # :52
    soil = (sm != 0) & ~lake                                                               # :51, :53, :56
    with np.errstate(invalid='ignore'):
        c2 = soil & (b >= sm)                                                              # :54, :57
        c3 = soil & (b < sm)                                                               # :55, :58
        q[lake] = np.maximum(0, p[lake] - pet[lake])                                       # :61
        q[np.isnan(q)] = 0.0                                                               # :62
        aet[lake] = np.minimum(p[lake], pet[lake])                                         # :63
    nanaet = np.isnan(aet)
    aet[nanaet] = pet[nanaet]                                                              # :64
    q[c2] = b[c2] - sm[c2]                                                                 # :67-69
    sav[c2] = sm[c2]
    aet[c2] = pet[c2]
    alpha = 1
    t3 = chstor[c3] + p[c3]                                                                # :73
    t5 = (5. * chstor[c3] / sm[c3] - 2. * (chstor[c3] / sm[c3]) ** 2.) / 3.                # :74
    t6 = np.minimum(np.ones_like(t5), t5)                                                  # :75
    t7 = pet[c3] * np.maximum(0.1 * np.ones_like(t6), t6)                                  # :76
    aet[c3] = np.minimum(t3, t7)                                                           # :77
    t8 = chstor[c3] * (1 - np.exp(-alpha * chstor[c3] / sm[c3])) / (1 - np.exp(-alpha)) + (p[c3] - aet[c3])   # :79
    sav[c3] = np.minimum(sm[c3], t8)                                                       # :80
    dry = c3 & (sav <= 0)                                                                  # :83
    sav[dry] = 0                                                                           # :84
    aet[dry] = p[dry] + chstor[dry]                                                        # :85
    q[c3] = np.maximum(np.zeros_like(q[c3]), chstor[c3] + p[c3] - aet[c3] - sav[c3])       # :86
    return aet, q, sav


def stepwise_run(temp, dtr, precip, lat_rad, sm_max, sm_prev, yr_imth_dys, runoff_spinup):
    """
    Hargreaves + GWAM as ConfigRunner.run drives them: spin-up pass over months 0..runoff_spinup-1
    (only `sm_prev` survives it), then all months.  Inputs [ncell, nmonths]; precip keeps its NaNs,
    temp and dtr go through nan_to_num (components.py:143-157).  Returns dict(pet, aet, q, sav,
    sm_after_spinup).
    """
    n, m = precip.shape
    solar_dec, dr = calc_sinusoidal_factor(yr_imth_dys)
    t_all, d_all = np.nan_to_num(temp), np.nan_to_num(dtr)
    pet = np.zeros((n, m))
    for k in range(m):
        pet[:, k] = hargreaves_pet(t_all[:, k], d_all[:, k], lat_rad, solar_dec[k], dr[k], yr_imth_dys[k, 2])
    sm = np.array(sm_prev, dtype=float)
    for k in range(runoff_spinup):
        _, _, sm = gwam_runoffgen(pet[:, k], precip[:, k], sm_max, sm)
    sm_spun = sm.copy()
    aet, q, sav = np.zeros((n, m)), np.zeros((n, m)), np.zeros((n, m))
    for k in range(m):
        aet[:, k], q[:, k], sav[:, k] = gwam_runoffgen(pet[:, k], precip[:, k], sm_max, sm)
        sm = sav[:, k].copy()
    return dict(pet=pet, aet=aet, q=q, sav=sav, sm_after_spinup=sm_spun)
