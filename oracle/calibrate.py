"""
Calibration-objective oracle (test infrastructure only).

Restates `basin_runoff` (xanthos/calibrate/calibrate_abcd.py:134-162, runoff
target only: the streamflow branch :164-173 is broken in the reference) and
`objective_kge` (:176-213).
"""

import numpy as np

from .abcd import abcd_emulate


def basin_series(pars, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas):
    """
    Modelled basin runoff series [n_months] for ONE parameter vector.

    pet/precip/tmin are the basin's rows [n_b, M]; parameters are repeated to
    every cell and all cells share basin id 0 (:143-149).
    """
    n_b = pet.shape[0]
    p = np.repeat(np.asarray(pars, dtype=float)[np.newaxis, ...], n_b, axis=0)
    if tmin is None and p.shape[1] == 4:
        p = np.concatenate([p, np.zeros((n_b, 1))], axis=1)
    _, q, _, _, _ = abcd_emulate(p, pet, precip, tmin, np.zeros(n_b), n_months, runoff_spinup)
    rsim = q.T                                                                 # [M, n_b]
    if obs_unit == 'km3_per_mth':
        return np.nansum(rsim * bsn_areas * 1e-6, 1)                           # :159
    elif obs_unit == 'mm_per_mth':
        return np.nansum(rsim, 1)                                              # :162
    raise ValueError(obs_unit)


def kge_distance(modelled, observed):
    """Euclidean distance from the KGE optimum (:197-211); KGE = 1 - distance."""
    sd_m, sd_o = np.std(modelled), np.std(observed)
    m_m, m_o = np.mean(modelled), np.mean(observed)
    relvar = sd_m / sd_o
    bias = m_m / m_o
    r = np.corrcoef(observed, modelled)[1, 0]
    return (((r - 1) ** 2) + ((relvar - 1) ** 2) + ((bias - 1) ** 2)) ** 0.5


def objective_kge(pars, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas, bsn_robs):
    """objective_kge(pars, basin_runoff, 0, ...) of the reference for the runoff target."""
    mod = basin_series(pars, pet, precip, tmin, n_months, runoff_spinup, obs_unit, bsn_areas)
    return kge_distance(mod, bsn_robs)
