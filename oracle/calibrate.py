"""
Calibration-objective oracle (test infrastructure only).

Restates `basin_runoff` (xanthos/calibrate/calibrate_abcd.py:134-162) and
`objective_kge` (:176-213) for the runoff target, and the INTENDED semantics of
the streamflow target (`set_calibrate = 1`, :164-173) - see
`objective_kge_streamflow`: the reference's own branch hands the whole
[ncell, nmonths] Avg_ChFlow array of Components.calculate_routing
(components.py:249-296) to np.std / np.mean / np.corrcoef, i.e. it takes the
moments over every cell of the globe and the correlation with the series of
global cell 0 (after building a 67,421 x 67,421 covariance matrix, 36 GB), so
it has no usable behaviour to pin; what is restated here is what
docs/calibration_tutorial.md describes - the routed flow at the basin's outlet
against the observed streamflow - and it is labelled as such everywhere.
PARITY UNPINNED for `objective_kge_streamflow` / `outlet_cells` (no reference output exists to pin them to;
tests/test_oracle.py::test_reference_streamflow_branch_has_no_usable_behaviour shows what the reference does);
the runoff-target functions above are pinned to the live reference (validate_against_reference.py).
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


def outlet_cells(basin_ids, dsid, area):
    """
    Outlet cell (0-based index) of every basin id 1 .. max: the cell of the basin with the largest
    drainage area (its own area plus that of every cell upstream of it), lowest index on ties.
    `dsid` is the 1-based downstream cell id per cell, 0 = none (routing/mrtm.py:85-120).
    INTENDED semantics of the streamflow target (module docstring): the reference names no cell.
    """
    basin_ids = np.asarray(basin_ids).astype(int)
    down = np.asarray(dsid).astype(np.int64) - 1
    n = len(down)
    acc = np.asarray(area, dtype=float).copy()
    indeg = np.zeros(n, dtype=int)
    for j in range(n):
        if down[j] >= 0:
            indeg[down[j]] += 1
    stack = [j for j in range(n) if indeg[j] == 0]
    while stack:                       # leaves first; a cell hands its total to its receiver once it is complete
        j = stack.pop()
        r = down[j]
        if r >= 0:
            acc[r] += acc[j]
            indeg[r] -= 1
            if indeg[r] == 0:
                stack.append(r)
    out = np.full(basin_ids.max(), -1, dtype=int)
    for b in range(1, basin_ids.max() + 1):
        cells = np.nonzero(basin_ids == b)[0]
        if len(cells):
            out[b - 1] = cells[np.argmax(acc[cells])]      # argmax returns the first maximum = lowest index
    return out


def objective_kge_streamflow(pars, pet, precip, tmin, n_months, runoff_spinup, basin_idx, arr_shp, route_fn, outlet,
                             bsn_qobs):
    """
    KGE distance between the routed streamflow at the basin's outlet and the observed streamflow (m3/s).

    As the reference (:164-171): the basin's ABCD runoff is put back into a global array of zeros and routed
    (`route_fn(rsim [ncell, nmonths]) -> Avg_ChFlow [ncell, nmonths]`, the router_func of Components.calibrate,
    components.py:486-497).  INTENDED part: the modelled series is row `outlet` of that array.
    """
    n_b = pet.shape[0]
    p = np.repeat(np.asarray(pars, dtype=float)[np.newaxis, ...], n_b, axis=0)
    if tmin is None and p.shape[1] == 4:
        p = np.concatenate([p, np.zeros((n_b, 1))], axis=1)
    _, q, _, _, _ = abcd_emulate(p, pet, precip, tmin, np.zeros(n_b), n_months, runoff_spinup)   # q: [n_b, M]
    rsim = np.zeros(shape=arr_shp)                                            # :169
    rsim[basin_idx, :] = q                                                    # :170 (np.put with he.rsim)
    avg = route_fn(rsim)
    return kge_distance(avg[outlet, :], bsn_qobs)
