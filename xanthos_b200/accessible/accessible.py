"""
Accessible water by basin - counterpart of xanthos/accessible/accessible.py.

`AccessibleWater(settings, ref, runoff)` keeps the reference's inputs, arithmetic and output file.  The two
O(ncell x nmonths) parts run on the device from the resident runoff field: mm/month -> km3/year
(xan_year_sum_scaled, accessible.py:34-39) and the basin aggregation (xan_group_sum, :41-51, same accumulation
order); what is left is [n_basins x n_years] host arithmetic restated from the reference (rolling mean, GCAM
years, base flow, environmental flow requirement, reservoir term - including the (nb,) + (nb, 1) broadcast of
`accessible_water`, :132-141, which makes the reservoir term the minimum capacity over all basins).
"""

import os
import logging

import numpy as np
import pandas as pd

from .. import _cuda as C
from ..diagnostics.time_series import group_sum_device


def basin_annual_runoff_device(runoff, area, basin_ids):
    """[ncell, nmonths] mm/month -> cuda tensor [NB, nyears] km3/year per basin."""
    torch = C.torch_cuda()
    f = C.as_field(runoff)
    ny = int(f.nmonths / 12)
    if ny < 1:
        raise C.ValidationException("AccessibleWater needs at least one whole year of runoff")
    conversion = C.dev_vector(np.asarray(area, dtype=np.float64) / 1e6)     # mm -> km3
    q = torch.empty((ny, f.ld), dtype=torch.float64, device='cuda')
    C.check(C.lib().xan_year_sum_scaled(C.ptr(f.t), C.ptr(conversion), f.ncell, f.nmonths, f.ld, C.ptr(q),
                                        C.stream_ptr()))
    return group_sum_device(basin_ids, q)


def AccessibleWater(settings, ref, runoff):
    """Accessible water per basin and GCAM year; writes accessible_water_km3peryr_<name>.csv and returns the table."""
    capacity = pd.read_csv(settings.ResCapacityFile, header=None, names=['res_capacity']).values      # (nb, 1)
    bfi = pd.read_csv(settings.BfiFile)['bfi_avg'].values

    annual = basin_annual_runoff_device(runoff, ref.area, ref.basin_ids).cpu().numpy()                # [nb, nyears]
    q_gcam = QInGCAMYears(RollingWindowFilter(annual, settings.MovingMeanWindow), settings)
    baseflow = (q_gcam.T * bfi).T                                                                      # accessible.py:57

    # environmental flow requirement: a share of the mean annual runoff of the historical years (:60-73)
    n_hist = annual.shape[1]
    if settings.StartYear > settings.HistEndYear:
        logging.warning('No historical data used in calculating Environmental Flow '
                        'Requirements (EFR) per basin for Accessible Water')
    elif settings.EndYear > settings.HistEndYear:
        n_hist = settings.HistEndYear - settings.StartYear + 1
    efr = settings.Env_FlowPercent * np.mean(annual[:, :n_hist], axis=1)

    ac = accessible_water(q_gcam, baseflow, efr, capacity)
    genGCAMOutput(os.path.join(settings.OutputFolder, 'accessible_water_km3peryr_{}.csv'.format(settings.OutputNameStr)),
                  ac, ref.basin_names, settings)
    return ac


def RollingWindowFilter(data, window, Dimension=0):
    """
    Centred moving average of width `window` along the rows (Dimension = 0) or columns (1) of a 2-D array (or along
    a 1-D array); the first and the last value are the plain means of the half window at that end (accessible.py:78-103).
    numpy.convolve does the sums, as in the reference, so the values are bit-identical.
    """
    data = np.asarray(data, dtype=float)
    kernel = np.full(window, 1.0 / window)
    if data.ndim == 1:
        return np.convolve(data, kernel, 'same')
    half = (window - 1) // 2 + 1
    rows = data if Dimension == 0 else data.T
    out = np.empty(rows.shape, dtype=float)
    for k, series in enumerate(rows):
        out[k] = np.convolve(series, kernel, 'same')
        out[k, 0] = series[:half].mean()
        out[k, -1] = series[len(series) - half:].mean()
    return out if Dimension == 0 else out.T


def _gcam_years(settings):
    return list(range(settings.GCAM_StartYear, settings.GCAM_EndYear + 1, settings.GCAM_YearStep))


def QInGCAMYears(qs, settings):
    """The columns of the GCAM target years out of the yearly table (accessible.py:106-116)."""
    cols = [y - settings.StartYear for y in _gcam_years(settings)]
    if min(cols) < 0 or max(cols) >= qs.shape[1]:
        raise ValueError("GCAM year outside {}-{}".format(settings.StartYear, settings.EndYear))
    return np.ascontiguousarray(qs[:, cols], dtype=float)


def accessible_water(qtot, base, efr, res):
    """
    min(total - EFR, baseflow - EFR + reservoir capacity), not below zero (accessible.py:119-129).  The reference adds
    the (nb, 1) capacity column to an (nb,) vector, which broadcasts to every pair of basins before the minimum is
    taken: the reservoir term of every basin is the SMALLEST capacity of all basins (x + r is monotone in r, so
    min_k (x + r_k) == x + min_k r_k exactly).  A 1-D capacity vector is applied basin by basin.
    """
    res = np.asarray(res, dtype=float)
    cap = res.min() if res.ndim == 2 else res.reshape(-1, 1)
    a = qtot - efr[:, None]
    b = (base - efr[:, None]) + cap
    c = np.minimum(a, b)
    return np.where(c < 0, 0.0, c)


def genGCAMOutput(filename, data, bdf, settings):
    """id, basin name and one column per GCAM year, values written as numpy prints them (accessible.py:132-146)."""
    table = {'id': np.arange(1, len(bdf) + 1).astype(str), 'name': np.asarray(bdf).astype(str)}
    text = np.asarray(data).astype(str)
    for k, year in enumerate(_gcam_years(settings)):
        table[str(year)] = text[:, k]
    pd.DataFrame(table).to_csv(filename, index=False)
