"""
ABCD runoff oracle (test infrastructure only).

Restates xanthos/runoff/abcd.py: the monthly recurrence `ABCD.abcd_dist`
(:171-228), the rain/snow split (:141-169), the spin-up re-initialisation
`set_vals` (:246-282) and the driver `abcd_execute` / `_run_basins`
(:314-422).  Arrays are [ncell, nmonths] like the reference's.
"""

import numpy as np

TRAIN = 2.5      # abcd.py:103
TSNOW = 0.6      # abcd.py:104
SW0 = 100.0      # abcd.py:82-84 (initial soil moisture)
GW0 = 500.0      # abcd.py:82-84 (initial groundwater)


def _pass(pars, pet, precip, tmin, steps, sw0, gw0, keep):
    """
    One pass of the recurrence over `steps` months for all cells at once.

    pars [n, 5]; pet/precip/tmin [n, >=steps] (tmin may be None); sw0/gw0 scalars
    or [n].  Returns dict of [steps, n] arrays named in `keep`.
    """
    n = pars.shape[0]
    a = pars[:, 0]
    b = pars[:, 1] * 1000                                                     # :48
    c = pars[:, 2]
    d = pars[:, 3]
    nosnow = tmin is None
    m = 0 if nosnow else pars[:, 4]
    a_times2 = a * 2                                                           # :54-56
    b_over_a = b / a
    d_plus_1 = d + 1

    out = {k: np.zeros((steps, n)) for k in keep}
    snowpack_prev = np.zeros(n)
    sw_prev = np.broadcast_to(np.asarray(sw0, dtype=float), (n,))
    g_prev = np.broadcast_to(np.asarray(gw0, dtype=float), (n,))

    for i in range(steps):
        p = precip[:, i]
        e = pet[:, i]
        if nosnow:
            rain = p
            snm = 0.0
        else:
            t = tmin[:, i]
            allrain = t > TRAIN
            mixed = (t <= TRAIN) & (t >= TSNOW)
            allsnow = t < TSNOW
            frac = (TRAIN - t) / (TRAIN - TSNOW)
            snow = np.zeros(n)                                                 # :141-169
            rain = np.zeros(n)
            snow[mixed] = (p * (TRAIN - t) / (TRAIN - TSNOW))[mixed]
            rain[allrain] = p[allrain]
            rain[mixed] = p[mixed] - snow[mixed]
            snow[allsnow] = p[allsnow]
            snowpack = (0 + snow) if i == 0 else (snowpack_prev + snow)        # :178-181
            snm = np.zeros(n)                                                  # :189-192
            snm[allrain] = (snowpack * m)[allrain]
            snm[mixed] = ((snowpack * m) * frac)[mixed]
            snowpack = snowpack - snm                                          # :195
            snowpack_prev = snowpack

        if i == 0:                                                             # :198-201
            w = rain + sw_prev
        else:
            w = rain + sw_prev + snm

        rpt = w + b                                                            # :204-206
        x = rpt / a_times2
        y = x - np.sqrt(np.square(x) - (w * b_over_a))
        sw = y * np.exp(-e / b)                                                # :209
        awet = w - y                                                           # :212-213
        c_awet = c * awet
        g = (g_prev + c_awet) / d_plus_1                                       # :216-219
        aet = y - sw                                                           # :222-224
        aet = np.maximum(0, aet)
        aet = np.minimum(e, aet)
        sw = y - aet                                                           # :225
        q = (awet - c_awet) + d * g                                            # :226

        sw_prev, g_prev = sw, g
        if 'aet' in out:
            out['aet'][i] = aet
        if 'q' in out:
            out['q'][i] = q
        if 'sw' in out:
            out['sw'][i] = sw
        if 'g' in out:
            out['g'][i] = g
    return out


def _basin_reinit(sw, g, basin_ids):
    """Mean over the last three Decembers of the per-basin nanmean (set_vals, :246-282)."""
    if sw.shape[0] < 25:
        raise IndexError("index -25 is out of bounds for axis 0 with size {}".format(sw.shape[0]))
    dec = [-1, -13, -25]
    sm_r = sw[dec, :]
    gs_r = g[dec, :]
    sm0 = np.empty(basin_ids.shape)
    gs0 = np.empty(basin_ids.shape)
    with np.errstate(invalid='ignore'), np.testing.suppress_warnings() as sup:
        sup.filter(RuntimeWarning)
        for b in np.unique(basin_ids):
            idx = (b == basin_ids)
            sm0[idx] = np.mean(np.nanmean(sm_r[:, idx], axis=1))
            gs0[idx] = np.mean(np.nanmean(gs_r[:, idx], axis=1))
    return sm0, gs0


def abcd_emulate(pars, pet, precip, tmin, basin_ids, n_months, spinup_steps):
    """
    ABCD.emulate() (:305-311) for an arbitrary set of cells.
    Returns aet, q, sw as [n, n_months] plus the re-initialised (sw0, gw0) [n].
    """
    spin = _pass(pars, pet, precip, tmin, spinup_steps, SW0, GW0, keep=('sw', 'g'))
    sw0, gw0 = _basin_reinit(spin['sw'], spin['g'], basin_ids)
    sim = _pass(pars, pet, precip, tmin, n_months, sw0, gw0, keep=('aet', 'q', 'sw'))
    return sim['aet'].T, sim['q'].T, sim['sw'].T, sw0, gw0


def abcd_execute(n_basins, basin_ids, pet, precip, tmin, pars_by_basin, n_months, spinup_steps):
    """
    abcd_execute (:394-422) with the calibration file already loaded as
    `pars_by_basin` [n_basins, 5].  Cells whose basin id lies outside
    min_id .. min_id + n_basins - 1 are returned as NaN (the reference leaves
    them uninitialised, :384-389).
    """
    basin_ids = np.asarray(basin_ids).astype(int)
    n = basin_ids.shape[0]
    min_b = basin_ids.min()
    sel = (basin_ids >= min_b) & (basin_ids < min_b + n_basins)
    idx = np.nonzero(sel)[0]
    pars = pars_by_basin[basin_ids - 1][idx]                                   # :332-333
    tm = None if tmin is None else tmin[idx]
    aet, q, sw, _, _ = abcd_emulate(pars, pet[idx], precip[idx], tm, basin_ids[idx], n_months, spinup_steps)
    out = [np.full((n, n_months), np.nan) for _ in range(4)]
    out[0][idx] = pet[idx][:, :n_months]
    out[1][idx] = aet
    out[2][idx] = q
    out[3][idx] = sw
    return tuple(out)
