"""
ORACLE (test infrastructure, not a product path): numpy restatement of the post-processing scans of
SURVEY.md section 8 row f3.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Pinned bitwise to the reference executed in the build container
(`python -m oracle.validate_against_reference`, section "postproc"): DroughtStats.getthresh / droughtstats
(xanthos/drought/drought_stats.py), Aggregation_Map (xanthos/diagnostics/time_series.py:126-138) and the
AccessibleWater chain (xanthos/accessible/accessible.py).
"""

import numpy as np


# ---- drought (drought_stats.py) -----------------------------------------------------------------------------
def virtual_index(n, q):
    """numpy's index of the q-quantile among n sorted samples, method 'linear':
    numpy/lib/_function_base_impl.py::_QuantileMethods['linear'] = (n - 1) * q, split into floor and fraction."""
    vi = (n - 1) * q
    prev = int(np.floor(vi))
    return prev, vi - prev


def lerp(a, b, t):
    """numpy's two-sided linear interpolation (_lerp): a + (b - a) t, or b - (b - a)(1 - t) when t >= 0.5."""
    diff = b - a
    return np.where(t >= 0.5, b - diff * (1 - t), a + diff * t)


def getthresh(histout, nper, quantile=0.1):
    """drought_stats.py:150-171: np.percentile(histout.reshape(nyear, nper, ngrid), quantile * 100, axis=0)."""
    ntime, ngrid = histout.shape
    nyear = int(ntime / nper)
    h = np.sort(np.reshape(histout, (nyear, nper, ngrid)), axis=0)       # NaN sorts last
    q = np.true_divide(quantile * 100, 100)                               # percentile -> quantile, as numpy does
    prev, gamma = virtual_index(nyear, q)
    nxt = min(prev + 1, nyear - 1)
    out = lerp(h[prev], h[nxt], gamma)
    out[np.isnan(h[-1])] = np.nan
    return out


def calculate_thresholds(histout, start_year, threshold_start_year, threshold_end_year, nper):
    """drought_stats.py:67-83 (the end index is (eyear + 1 - syear) * 12, not an offset from smonth: kept)."""
    smonth = (threshold_start_year - start_year) * 12
    emonth = (threshold_end_year + 1 - threshold_start_year) * 12
    return getthresh(histout[smonth:emonth, :], nper)


def droughtstats(hydroout, threshvals):
    """drought_stats.py:85-148, time loop restated per step ([ntime, ngrid] in, three of the same out)."""
    ntime, nthresh = hydroout.shape[0], threshvals.shape[0]
    S = np.empty_like(hydroout)
    I = np.empty_like(hydroout)
    D = np.empty_like(hydroout)
    with np.errstate(invalid='ignore', divide='ignore'):
        dry = hydroout[0] < threshvals[0]
        D[0] = np.where(dry, 1.0, 0.0)
        I[0] = S[0] = np.where(dry, (threshvals[0] - hydroout[0]) / threshvals[0], 0.0)
        for t in range(1, ntime):
            th, h = threshvals[t % nthresh], hydroout[t]
            dry = h < th
            D[t] = np.where(dry, D[t - 1] + 1, 0.0)
            S[t] = np.where(dry, S[t - 1] + (th - h) / th, 0.0)
            I[t] = np.where(dry, S[t] / D[t], 0.0)
    return S, I, D


# ---- group sums (time_series.py:126-138, accessible.py:41-51) ---------------------------------------------------
def aggregation_map(id_map, values):
    """out[g - 1, t] = sum in ascending cell index of the non-NaN values[cell, t] with id_map[cell] == g > 0."""
    id_map = np.asarray(id_map).astype(int)
    nb = int(id_map.max())
    out = np.zeros((nb, values.shape[1]))
    for c in range(values.shape[0]):       # cell order = the reference's accumulation order
        g = id_map[c]
        if g > 0:
            ok = ~np.isnan(values[c])
            out[g - 1, ok] += values[c, ok]
    return out


# ---- accessible water (accessible.py) ------------------------------------------------------------------------------
def yearly_km3(runoff, area):
    """accessible.py:34-39: np.sum over 12 months (numpy's pairwise order) times area / 1e6."""
    ny = int(runoff.shape[1] / 12)
    conversion = area / 1e6
    q = np.zeros((runoff.shape[0], ny))
    for i in range(ny):
        q[:, i] = np.sum(runoff[:, i * 12:(i + 1) * 12], axis=1) * conversion
    return q


def rolling_window_filter(data, window):
    """accessible.py:78-103, Dimension = 0 (per row)."""
    weights = np.repeat(1.0, window) / window
    it = int((window - 1) / 2) + 1
    sma = np.zeros(data.shape)
    for i in range(data.shape[0]):
        sma[i, :] = np.convolve(data[i, :], weights, 'same')
        sma[i, 0] = np.mean(data[i, :it])
        sma[i, data.shape[1] - 1] = np.mean(data[i, data.shape[1] - it:])
    return sma


def accessible_water_chain(map_runoff, start_year, end_year, hist_end_year, gcam_years, window, bfi, res_capacity,
                           env_flow_percent):
    """accessible.py:53-76 and :106-130 from the basin-aggregated annual runoff on."""
    valid = list(range(start_year, end_year + 1))
    qs = rolling_window_filter(map_runoff, window)
    q_gcam = np.stack([qs[:, valid.index(y)] for y in gcam_years], axis=1)
    bflow = np.transpose(np.transpose(q_gcam) * np.asarray(bfi))
    if start_year > hist_end_year or end_year <= hist_end_year:
        efr = env_flow_percent * np.mean(map_runoff, axis=1)
    else:
        efr = env_flow_percent * np.mean(map_runoff[:, :valid.index(hist_end_year) + 1], axis=1)
    res = np.asarray(res_capacity).reshape(-1, 1)
    ac = np.zeros(q_gcam.shape)
    for i in range(q_gcam.shape[1]):
        a = q_gcam[:, i] - efr
        b = bflow[:, i] - efr + res                # (nb,) + (nb, 1) broadcasts to (nb, nb), as in the reference
        c = np.min(np.vstack((a, b)), axis=0)
        ac[:, i] = np.where(c < 0, 0, c)
    return ac


def agg_to_year(arr, func='sum'):
    """
    OutWriter.agg_to_year (data_writer/out_writer.py:237-248): `df.groupby(np.arange(ncol) // 12, axis=1).agg(func)`.
    The axis=1 form no longer exists in pandas 3; grouping the transposed frame is pandas' documented replacement
    and keeps its semantics (NaN skipped: an all-NaN year sums to 0.0 and averages to NaN; Kahan summation).
    """
    import pandas as pd
    df = pd.DataFrame(np.asarray(arr))
    return df.T.groupby(np.arange(df.shape[1]) // 12).agg(func).T.values


def agg_spatial(arr, id_map):
    """
    OutWriter.agg_spatial (out_writer.py:250-265) without the name join: `df.groupby('id').sum()` -> (ids present,
    [n_ids, ntime]); NaN skipped.
    """
    import pandas as pd
    df = pd.DataFrame(np.asarray(arr))
    df['id'] = np.asarray(id_map)
    g = df.groupby('id', as_index=False).sum()
    return g['id'].values.astype(int), g.drop(columns='id').values
