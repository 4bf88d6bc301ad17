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
    """Calculate accessible water per basin."""
    bdf = ref.basin_names
    rdf = pd.read_csv(settings.ResCapacityFile, header=None, names=['res_capacity'])
    bfi = pd.read_csv(settings.BfiFile)['bfi_avg']

    Map_runoff = basin_annual_runoff_device(runoff, ref.area, ref.basin_ids).cpu().numpy()

    qs = RollingWindowFilter(Map_runoff, settings.MovingMeanWindow)
    q_gcam = QInGCAMYears(qs, settings)
    bflow = np.transpose(np.transpose(q_gcam) * np.array(bfi))

    if settings.StartYear > settings.HistEndYear:
        logging.warning('No historical data used in calculating Environmental Flow '
                        'Requirements (EFR) per basin for Accessible Water')
        edf = settings.Env_FlowPercent * np.mean(Map_runoff, axis=1)
    elif settings.EndYear <= settings.HistEndYear:
        edf = settings.Env_FlowPercent * np.mean(Map_runoff, axis=1)
    else:
        hey = list(range(settings.StartYear, settings.EndYear + 1)).index(settings.HistEndYear)
        edf = settings.Env_FlowPercent * np.mean(Map_runoff[:, :(hey + 1)], axis=1)

    ac = accessible_water(q_gcam, bflow, edf, rdf.values)
    filename = os.path.join(settings.OutputFolder, 'accessible_water_km3peryr_{}.csv'.format(settings.OutputNameStr))
    genGCAMOutput(filename, ac, bdf, settings)
    return ac


def RollingWindowFilter(data, window, Dimension=0):
    """Centred moving average with shortened windows at both ends (accessible.py:78-103)."""
    weights = np.repeat(1.0, window) / window
    it = int((window - 1) / 2) + 1
    if data.ndim == 1:
        return np.convolve(data, weights, 'same')
    sma = np.zeros(data.shape, dtype=float)
    if Dimension == 1:
        for i in range(data.shape[1]):
            sma[:, i] = np.convolve(data[:, i], weights, 'same')
            sma[0, i] = np.mean(data[:it, i])
            sma[data.shape[0] - 1, i] = np.mean(data[data.shape[0] - it:, i])
    elif Dimension == 0:
        for i in range(data.shape[0]):
            sma[i, :] = np.convolve(data[i, :], weights, 'same')
            sma[i, 0] = np.mean(data[i, :it])
            sma[i, data.shape[1] - 1] = np.mean(data[i, data.shape[1] - it:])
    return sma


def QInGCAMYears(qs, settings):
    """Columns of the GCAM target years (accessible.py:106-116)."""
    valid = list(range(settings.StartYear, settings.EndYear + 1))
    gcam = list(range(settings.GCAM_StartYear, settings.GCAM_EndYear + 1, settings.GCAM_YearStep))
    q_gcam = np.zeros((qs.shape[0], len(gcam)), dtype=float)
    for i, y in enumerate(gcam):
        q_gcam[:, i] = qs[:, valid.index(y)]
    return q_gcam


def accessible_water(qtot, base, efr, res):
    """min(qtot - efr, base - efr + res) clipped at 0 (accessible.py:119-129, broadcasting kept)."""
    ac = np.zeros(qtot.shape, dtype=float)
    for i in range(qtot.shape[1]):
        a = qtot[:, i] - efr
        b = base[:, i] - efr + res
        c = np.min(np.vstack((a, b)), axis=0)
        ac[:, i] = np.where(c < 0, 0, c)
    return ac


def genGCAMOutput(filename, data, bdf, settings):
    """id, name and accessible water by GCAM year as .csv (accessible.py:132-146)."""
    years = list(map(str, range(settings.GCAM_StartYear, settings.GCAM_EndYear + 1, settings.GCAM_YearStep)))
    hdr = "id,name," + ",".join(years)
    MapId = np.arange(1, len(bdf) + 1, 1, dtype=int).astype(str)
    newdata = np.insert(data.astype(str), 0, bdf, axis=1)
    Result = np.insert(newdata.astype(str), 0, MapId, axis=1)
    df = pd.DataFrame(Result)
    df.columns = hdr.split(',')
    df.to_csv(filename, index=False)
