"""
Deterministic synthetic worlds and forcing for parity tests and benchmarks.

The Zenodo example data of the reference is not available offline, so every
test and benchmark of this repository runs on a seeded synthetic world that has
the same shapes and file-level conventions as the reference inputs
(reference: xanthos/data_reader/data_load.py:47-72 for the static grids,
:92-135 for the Penman-Monteith tables, :200-211 for the routing vectors).

Host-side numpy only: this module builds *inputs*; it never computes a model
result.  The same arrays are handed to the CUDA path and to the oracle.

Conventions (reference xanthos/routing/mrtm.py:88-95, 237-240):
  * coords[:, 0..4] = id (1-based), lon, lat, ilon (1-based column), ilat
    (1-based row, south -> north);
  * D8 codes 1=E 2=SE 4=S 8=SW 16=W 32=NW 64=N 128=NE, "N" = +1 grid row,
    -9999 = missing, 0 = no direction.
"""

import heapq
from dataclasses import dataclass, field

import numpy as np
from scipy import ndimage

# (drow, dcol) -> D8 code, orientation of xanthos/routing/mrtm.py:237-240
D8_CODE = {(0, 1): 1, (-1, 1): 2, (-1, 0): 4, (-1, -1): 8,
           (0, -1): 16, (1, -1): 32, (1, 0): 64, (1, 1): 128}


@dataclass
class World:
    """Static grids of one synthetic world (all host numpy, reference layouts)."""

    nrow: int
    ncol: int
    ncell: int
    n_basins: int
    coords: np.ndarray        # [N, 5] float64
    flow_dir: np.ndarray      # [N] float64 D8 codes (-9999 missing)
    flow_dist: np.ndarray     # [N] m, already floored at 1000 like the loader
    velocity: np.ndarray      # [N] m/s, already floored at 0
    area: np.ndarray          # [N] km2
    basin_ids: np.ndarray     # [N] int, 1..n_basins
    elev: np.ndarray          # [N, 1] m (Penman-Monteith pressure term)
    seed: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def lat(self):
        return self.coords[:, 2]

    @property
    def lat_radians(self):
        return np.radians(self.coords[:, 2])

    def settings(self):
        """Minimal settings bag for routing_mod.downstream/upstream."""
        from types import SimpleNamespace
        return SimpleNamespace(ngridrow=self.nrow, ngridcol=self.ncol, ncell=self.ncell)


def _smooth_field(rng, nrow, ncol, octaves=((6, 12, 1.0), (12, 24, 0.5), (24, 48, 0.25), (60, 120, 0.08))):
    """Sum of bicubically upsampled coarse noise grids, periodic-free."""
    out = np.zeros((nrow, ncol))
    for (r, c, amp) in octaves:
        r = max(2, min(r, nrow))
        c = max(2, min(c, ncol))
        coarse = rng.standard_normal((r, c))
        up = ndimage.zoom(coarse, (nrow / r, ncol / c), order=3, mode='nearest', grid_mode=True)
        out += amp * up[:nrow, :ncol]
    return out


def _priority_flood(land, elev):
    """
    Drainage directions for every land cell by priority-flood from the ocean.

    Returns (drow, dcol) int arrays on the full grid; every land cell points to
    the neighbour it was flooded from, coastal seeds point to an ocean (or
    off-grid) neighbour.  The result is a forest whose roots touch the ocean,
    i.e. a DAG by construction.
    """
    nrow, ncol = land.shape
    drow = np.zeros((nrow, ncol), dtype=np.int8)
    dcol = np.zeros((nrow, ncol), dtype=np.int8)
    done = ~land
    heap = []
    nbrs = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]

    rr, cc = np.nonzero(land)
    for r, c in zip(rr.tolist(), cc.tolist()):
        best = None
        for dr, dc in nbrs:
            r2, c2 = r + dr, c + dc
            off = (r2 < 0) or (r2 >= nrow) or (c2 < 0) or (c2 >= ncol)
            if off or not land[r2, c2]:
                # prefer a deterministic first hit; off-grid columns exercise the
                # longitude "wrap" arithmetic of mrtm.py:101-102
                best = (dr, dc)
                if off and (c2 < 0 or c2 >= ncol) and dr == 0:
                    break
        if best is not None:
            drow[r, c], dcol[r, c] = best
            done[r, c] = True
            heapq.heappush(heap, (float(elev[r, c]), r, c))

    filled = elev.astype(float).copy()
    while heap:
        h, r, c = heapq.heappop(heap)
        for dr, dc in nbrs:
            r2, c2 = r + dr, c + dc
            if r2 < 0 or r2 >= nrow or c2 < 0 or c2 >= ncol or done[r2, c2]:
                continue
            done[r2, c2] = True
            drow[r2, c2], dcol[r2, c2] = -dr, -dc
            filled[r2, c2] = max(filled[r2, c2], h + 1e-3)
            heapq.heappush(heap, (float(filled[r2, c2]), r2, c2))
    return drow.astype(int), dcol.astype(int)


def _flow_components(ncell, dsid):
    """Outlet index (0-based) reached by every cell; dsid is 1-based, -1 = outlet."""
    nxt = np.where(dsid > 0, dsid - 1, np.arange(ncell))
    root = nxt.copy()
    for _ in range(64):
        new = root[root]
        if np.array_equal(new, root):
            break
        root = new
    return root


def make_world(nrow=360, ncol=720, ncell=67420, n_basins=235, seed=0, edge_cases=True,
               cut_basins=False, coast_pull=3.0):
    """
    Build a synthetic world.

    :param coast_pull:  metres of terrain height per cell of distance to the coast;
                        small values give few, very large river trees.
    :param cut_basins:  if True, basin ids deliberately cut across flow
                        components (adversarial variant used to test the
                        basin/flow-graph consistency check).
    """
    rng = np.random.default_rng(seed)
    cap = max(1, nrow // 12)

    f = _smooth_field(rng, nrow, ncol)
    f[:cap, :] = -np.inf
    f[nrow - cap:, :] = -np.inf
    flat = f.ravel()
    if ncell > np.isfinite(flat).sum():
        raise ValueError("ncell larger than the admissible grid")
    order = np.argsort(-flat, kind='stable')
    land = np.zeros(nrow * ncol, dtype=bool)
    land[order[:ncell]] = True
    land = land.reshape(nrow, ncol)

    # ids: column (longitude) major, then south -> north
    cols, rows = np.nonzero(land.T)
    ids = np.arange(1, ncell + 1)
    lon = -180.0 + (cols + 0.5) * (360.0 / ncol)
    lat = -90.0 + (rows + 0.5) * (180.0 / nrow)
    coords = np.stack([ids, lon, lat, cols + 1, rows + 1], axis=1).astype(float)

    # terrain: height above the coast grows inland, plus smooth relief -> convergent rivers
    dist = ndimage.distance_transform_edt(land)
    relief = _smooth_field(rng, nrow, ncol, octaves=((9, 18, 1.0), (30, 60, 0.5), (90, 180, 0.25)))
    elev_grid = coast_pull * dist + 250.0 * relief
    drow, dcol = _priority_flood(land, elev_grid)

    code = np.zeros(ncell)
    dr = drow[rows, cols]
    dc = dcol[rows, cols]
    for (a, b), v in D8_CODE.items():
        code[(dr == a) & (dc == b)] = v

    if edge_cases and ncell > 50:
        pick = rng.choice(ncell, size=max(3, ncell // 400), replace=False)
        third = len(pick) // 3
        code[pick[:third]] = -9999.0          # missing -> self -> outlet (mrtm.py:243-245)
        code[pick[third:2 * third]] = 0.0     # no direction -> outlet

    # routing vectors (already clamped like data_load.py:204-208)
    flow_dist = rng.uniform(25e3, 70e3, ncell)
    velocity = rng.uniform(0.1, 2.0, ncell)
    if edge_cases and ncell > 50:
        flow_dist[rng.choice(ncell, size=max(2, ncell // 200), replace=False)] = 1000.0
        velocity[rng.choice(ncell, size=max(2, ncell // 300), replace=False)] = 0.0

    area = 3091.0 * np.cos(np.radians(lat)) * (0.5 * 360.0 / nrow) * (0.5 * 720.0 / ncol) * 4.0
    elev = np.clip(elev_grid[rows, cols], 0.0, None)
    elev = (3000.0 * elev / max(elev.max(), 1.0))[:, None]

    # basins = unions of flow components
    dsid = reference_free_downstream(coords, code, nrow, ncol)
    root = _flow_components(ncell, dsid)
    uniq, inv, counts = np.unique(root, return_inverse=True, return_counts=True)
    n_basins = int(min(n_basins, len(uniq)))
    big = np.argsort(-counts, kind='stable')[:n_basins]
    comp_basin = np.full(len(uniq), -1, dtype=int)
    comp_basin[big] = np.arange(1, n_basins + 1)
    # small components join the basin whose seed outlet is nearest
    seed_r = rows[uniq[big]].astype(float)
    seed_c = cols[uniq[big]].astype(float)
    rest = np.nonzero(comp_basin < 0)[0]
    if len(rest):
        rr_ = rows[uniq[rest]].astype(float)[:, None]
        cc_ = cols[uniq[rest]].astype(float)[:, None]
        d2 = (rr_ - seed_r[None, :]) ** 2 + (cc_ - seed_c[None, :]) ** 2
        comp_basin[rest] = np.argmin(d2, axis=1) + 1
    basin_ids = comp_basin[inv].astype(int)

    if cut_basins:
        # reassign by longitude stripes: basins now cut through river trees
        basin_ids = (np.floor((cols / ncol) * n_basins).astype(int) % n_basins) + 1

    return World(nrow=nrow, ncol=ncol, ncell=ncell, n_basins=n_basins, coords=coords,
                 flow_dir=code, flow_dist=flow_dist, velocity=velocity, area=area,
                 basin_ids=basin_ids, elev=elev, seed=seed,
                 meta={'n_components': int(len(uniq)), 'max_component': int(counts.max())})


def reference_free_downstream(coords, flow_dir, nrow, ncol):
    """
    Downstream id per cell, used only to *build* synthetic basins.

    Same arithmetic as the routing topology (xanthos/routing/mrtm.py:85-120) but
    kept private to the generator: model code never calls it.
    """
    n = coords.shape[0]
    ilat = coords[:, 4].astype(int) - 1
    ilon = coords[:, 3].astype(int) - 1
    grid = np.zeros((nrow, ncol), dtype=int)
    grid[ilat, ilon] = coords[:, 0].astype(int)
    fd = np.where(flow_dir == -9999.0, 0, flow_dir).astype(int)
    dlat = np.zeros(n, dtype=int)
    dlon = np.zeros(n, dtype=int)
    dlat[(fd & (2 + 4 + 8)) != 0] = -1
    dlat[(fd & (32 + 64 + 128)) != 0] = 1
    dlon[(fd & (1 + 2 + 128)) != 0] = 1
    dlon[(fd & (8 + 16 + 32)) != 0] = -1
    tlat = ilat + dlat
    tlon = ilon + dlon
    bad = (tlon < 0) | (tlon > ncol - 1)
    tlon[bad] = np.mod(tlon[bad] + 1, ncol)
    bad = (tlat < 0) | (tlat > nrow - 1)
    tlat[bad] = ilat[bad]
    tlon[bad] = ilon[bad]
    ds = grid[tlat, tlon]
    ds[(ds == 0) | (ds == coords[:, 0].astype(int))] = -1
    return ds


# --------------------------------------------------------------------------
# forcing
# --------------------------------------------------------------------------
def _seasonal_temperature(rng, world, nmonths):
    lat = world.lat
    base = 28.0 - 0.55 * np.abs(lat)                      # warm tropics, cold poles
    amp = 2.0 + 0.28 * np.abs(lat)
    phase = np.where(lat >= 0, 0.0, np.pi)
    mth = np.arange(nmonths)
    season = -np.cos(2 * np.pi * (mth[None, :] % 12) / 12.0 + phase[:, None])
    return base[:, None] + amp[:, None] * season + rng.normal(0.0, 2.5, (world.ncell, nmonths))


def pm_inputs(world, start_yr, end_yr, nlcs=8, lc_years=(1970, 1975, 1980, 1985, 1990, 1995, 2000), seed=1):
    """
    Penman-Monteith inputs with the attribute names DataLoader exposes
    (xanthos/data_reader/data_load.py:92-135); distributions per SURVEY.md §8(d).
    """
    rng = np.random.default_rng(seed)
    n = world.ncell
    m = (end_yr - start_yr + 1) * 12
    tair = _seasonal_temperature(rng, world, m)
    tmin = tair - rng.uniform(2.0, 8.0, (n, m))
    rhs = rng.uniform(20.0, 100.0, (n, m))
    # humidity break points of calc_fwet / calc_rh (penman_monteith.py:165-172, 205-209)
    k = max(1, (n * m) // 500)
    flat = rhs.ravel()
    for val in (70.0, 80.0, 90.0, 95.0, 99.99995, 100.0, 69.999999):
        flat[rng.choice(n * m, size=k, replace=False)] = val
    wind = rng.uniform(0.5, 8.0, (n, m))
    rsds = rng.uniform(20.0, 350.0, (n, m))
    rlds = rng.uniform(150.0, 450.0, (n, m))

    lct = rng.dirichlet(np.ones(nlcs), size=(n, len(lc_years))) * 100.0   # [n, years, nlcs]
    lct = np.ascontiguousarray(np.swapaxes(lct, 1, 2))                    # [n, nlcs, years]
    zero_rows = rng.choice(n, size=max(1, n // 300), replace=False)
    lct[zero_rows, :, :] = 0.0                                            # totpct == 0 -> 0.01

    par = {
        'cL': rng.uniform(0.001, 0.01, nlcs),
        'beta': rng.uniform(100.0, 300.0, nlcs),
        'rslimit': rng.uniform(1000.0, 5000.0, nlcs),
        'Tminopen': rng.uniform(8.0, 12.0, nlcs),
        'Tminclose': rng.uniform(-8.0, -6.0, nlcs),
        'VPDclose': rng.uniform(28.0, 45.0, nlcs),
        'VPDopen': rng.uniform(6.0, 10.0, nlcs),
        'RBLmin': rng.uniform(55.0, 70.0, nlcs),
        'RBLmax': rng.uniform(90.0, 100.0, nlcs),
        'rc': rng.uniform(0.005, 0.02, nlcs),
        'emiss': rng.uniform(0.9, 0.99, nlcs),
    }
    alpha = rng.uniform(0.05, 0.4, (nlcs, 12))
    lai = rng.uniform(0.0, 6.0, (nlcs, 12))
    lai[rng.integers(0, nlcs, 4), rng.integers(0, 12, 4)] = 0.0
    lai[min(3, nlcs - 1), :] = 0.0                                       # a bare class: fc_denom == 0 branch
    laimin = np.repeat(lai.min(axis=1, keepdims=True), 12, axis=1)
    laimax = np.repeat(lai.max(axis=1, keepdims=True), 12, axis=1)

    tairprev = np.zeros_like(tair)
    tairprev[1:, :] = tair[:-1, :]                                       # data_load.py:128-129 (cell shift)

    out = dict(tair_load=tair, TMIN_load=tmin, rhs_load=rhs, wind_load=wind, rsds_load=rsds,
               rlds_load=rlds, tairprev_load=tairprev, lct_load=lct, elev=world.elev.copy(),
               alpha=alpha, lai=lai, laimin=laimin, laimax=laimax,
               nlcs=nlcs, lc_years=list(lc_years), water_idx=0, snow_idx=6 if nlcs > 6 else nlcs - 1)
    out.update(par)
    return out


def hs_inputs(world, start_yr, end_yr, seed=2):
    """Hargreaves-Samani inputs hs_tas/hs_tmax/hs_tmin (data_load.py:88-90); NaN kept."""
    rng = np.random.default_rng(seed)
    n = world.ncell
    m = (end_yr - start_yr + 1) * 12
    tas = _seasonal_temperature(rng, world, m)
    tmax = tas + rng.uniform(2.0, 8.0, (n, m))
    tmin = tas - rng.uniform(2.0, 8.0, (n, m))
    k = max(1, (n * m) // 2000)
    tas.ravel()[rng.choice(n * m, size=k, replace=False)] = np.nan
    return dict(hs_tas=tas, hs_tmax=tmax, hs_tmin=tmin)


def thornthwaite_inputs(world, start_yr, end_yr, seed=5):
    rng = np.random.default_rng(seed)
    m = (end_yr - start_yr + 1) * 12
    tas = _seasonal_temperature(rng, world, m)
    k = max(1, tas.size // 2000)
    tas.ravel()[rng.choice(tas.size, size=k, replace=False)] = np.nan
    return dict(tair=tas)


def stepwise_inputs(world, start_yr, end_yr, seed=6):
    """
    Inputs of the step-wise (v1) path, Hargreaves PET + GWAM runoff: temp, dtr (with negatives and NaN),
    precip (NaN kept), maximum soil moisture (with 999 = water bodies, 0 = no soil) and the historic-mode
    initial soil moisture 0.5 * max (data_load.py:77-84, 147-182).
    """
    rng = np.random.default_rng(seed + 3000)
    n = world.ncell
    m = (end_yr - start_yr + 1) * 12
    temp = _seasonal_temperature(rng, world, m)
    dtr = rng.uniform(-1.0, 16.0, (n, m))
    k = max(1, (n * m) // 2000)
    temp.ravel()[rng.choice(n * m, size=k, replace=False)] = np.nan
    dtr.ravel()[rng.choice(n * m, size=k, replace=False)] = np.nan
    precip = np.abs(rng.normal(70.0, 60.0, (n, m)))
    precip.ravel()[rng.choice(n * m, size=k, replace=False)] = np.nan
    precip[rng.choice(n, size=max(1, n // 300), replace=False), :] = np.nan      # a cell without data
    sm_max = rng.uniform(15.0, 450.0, n)
    sm_max[rng.choice(n, size=max(2, n // 60), replace=False)] = 999.0           # lakes (data_load.py:154-160)
    sm_max[rng.choice(n, size=max(1, n // 150), replace=False)] = 0.0
    return dict(temp=temp, dtr=dtr, precip=precip, soil_moisture=sm_max, sm_prev=0.5 * sm_max)


def abcd_inputs(world, nmonths, seed=1, with_pet=True):
    """Precipitation (with ~0.1 % NaN), tmin, optional PET and [n_basins, 5] parameters."""
    rng = np.random.default_rng(seed + 1000)
    n = world.ncell
    precip = np.abs(rng.normal(80.0, 60.0, (n, nmonths)))
    nan_cells = rng.choice(n, size=max(1, n // 1000), replace=False)
    precip[nan_cells, rng.integers(0, nmonths, len(nan_cells))] = np.nan
    tmin = _seasonal_temperature(rng, world, nmonths) - 5.0
    # exact thresholds of set_rain_and_snow (abcd.py:141-169)
    tmin.ravel()[rng.choice(n * nmonths, size=8, replace=False)] = 2.5
    tmin.ravel()[rng.choice(n * nmonths, size=8, replace=False)] = 0.6
    lb = 1e-4
    pars = np.stack([rng.uniform(lb, 1 - lb, world.n_basins),
                     rng.uniform(lb, 8 - lb, world.n_basins),
                     rng.uniform(lb, 1 - lb, world.n_basins),
                     rng.uniform(lb, 1 - lb, world.n_basins),
                     rng.uniform(lb, 1 - lb, world.n_basins)], axis=1)
    pars[0, 0] = lb                                        # worst-case cancellation in sqrt term
    out = dict(precip=precip, tmin=tmin, pars=pars)
    if with_pet:
        out['pet'] = np.abs(rng.normal(90.0, 50.0, (n, nmonths)))
    return out


def runoff_input(world, nmonths, seed=3):
    """Runoff for routing-only runs: |N(50, 30)| with zero months."""
    rng = np.random.default_rng(seed + 2000)
    q = np.abs(rng.normal(50.0, 30.0, (world.ncell, nmonths)))
    q[:, rng.integers(0, nmonths, max(1, nmonths // 12))] = 0.0
    return q


def calibration_obs(basin_series, seed=4):
    """'VIC-like' observations: a model series times (1 + N(0, 0.05))."""
    rng = np.random.default_rng(seed + 3000)
    return basin_series * (1.0 + rng.normal(0.0, 0.05, basin_series.shape))


# --------------------------------------------------------------------------
# example project on disk (same file layout as the reference's example data)
# --------------------------------------------------------------------------
def write_example(root, world, start_yr, end_yr, pet='pm', routing=True, seed=1, runoff_spinup=None,
                  routing_spinup=None, calibrate=False, output_vars='q,avgchflow', project='synthetic',
                  postproc=None, project_overrides=None, extra_lines=None):
    """
    Write a synthetic project under `root` in the file formats the loader reads
    (xanthos/data_reader/ini_reader.py:428-435, data_load.py:47-72, 92-135, 200-211) and return the
    path of its .ini file plus the dict of in-memory inputs.  `postproc`: optional dict that switches the
    post-processing modules on, e.g. {'drought': {'drought_var': 'q', 'threshold_nper': 12, 'threshold_start_year': ..,
    'threshold_end_year': ..}, 'accessible_water': {'HistEndYear': .., 'GCAM_StartYear': .., 'GCAM_EndYear': ..,
    'GCAM_YearStep': .., 'MovingMeanWindow': .., 'Env_FlowPercent': ..}} (ini_reader.py:460-486).
    """
    import os
    n, m = world.ncell, (end_yr - start_yr + 1) * 12
    inp = os.path.join(root, 'input')
    ref = os.path.join(inp, 'reference')
    dirs = {k: os.path.join(inp, k) for k in ('pet', 'runoff', 'routing')}
    for d in [ref, os.path.join(dirs['pet'], pet), os.path.join(dirs['runoff'], 'abcd'),
              os.path.join(dirs['routing'], 'mrtm'), os.path.join(root, 'output')]:
        os.makedirs(d, exist_ok=True)
    np.savetxt(os.path.join(ref, 'Grid_Areas_ID.csv'), world.area * 100.0, fmt='%.17g')          # ha
    np.savetxt(os.path.join(ref, 'coordinates.csv'), world.coords, delimiter=',', fmt='%.17g')
    with open(os.path.join(ref, 'basin.csv'), 'w') as f:
        f.write('basin_id\n')
        np.savetxt(f, world.basin_ids, fmt='%d')
    with open(os.path.join(ref, 'BasinNames235.txt'), 'w') as f:
        f.write('\n'.join('Basin_{}'.format(i + 1) for i in range(world.n_basins)) + '\n')
    # country / GCAM region maps and their name files (data_load.py:60-69, 275-286): countries 0..nc (0 = none),
    # regions 1..nr; both follow the basins so that they are spatially coherent
    nc, nr = max(2, world.n_basins // 2), max(2, world.n_basins // 3)
    rng_ids = np.random.default_rng(seed + 4242)
    country = (world.basin_ids % (nc + 1)).astype(int)
    country[rng_ids.random(n) < 0.01] = 0
    region = (world.basin_ids % nr).astype(int) + 1
    with open(os.path.join(ref, 'country.csv'), 'w') as f:
        f.write('country_id\n')
        np.savetxt(f, country, fmt='%d')
    with open(os.path.join(ref, 'region32_grids.csv'), 'w') as f:
        f.write('region_id\n')
        np.savetxt(f, region, fmt='%d')
    with open(os.path.join(ref, 'country-names.csv'), 'w') as f:
        f.write('\n'.join('{},Country_{}'.format(i, i) for i in range(nc + 1)) + '\n')
    with open(os.path.join(ref, 'Rgn32Names.csv'), 'w') as f:
        f.write('region,region_id\n' + '\n'.join('Region_{},{}'.format(i, i) for i in range(1, nr + 1)) + '\n')

    data = {}
    pdir = os.path.join(dirs['pet'], pet)
    lines = ['[PET]', 'pet_module = {}'.format(pet)]
    if pet == 'pm':
        pm = pm_inputs(world, start_yr, end_yr, seed=seed)
        data.update(pm)
        for key, fn in (('tair_load', 'tas'), ('TMIN_load', 'tmin'), ('rhs_load', 'rhs'), ('wind_load', 'wind'),
                        ('rsds_load', 'rsds'), ('rlds_load', 'rlds')):
            np.save(os.path.join(pdir, fn + '.npy'), pm[key])
        np.save(os.path.join(pdir, 'lct.npy'), pm['lct_load'])
        np.save(os.path.join(pdir, 'elev.npy'), pm['elev'])
        et = np.zeros((pm['nlcs'], 13))
        for k, name in enumerate(('cL', 'beta', 'rslimit', None, None, 'Tminopen', 'Tminclose', 'VPDclose', 'VPDopen',
                                  'RBLmin', 'RBLmax', 'rc', 'emiss')):
            if name:
                et[:, k] = pm[name]
        np.savetxt(os.path.join(pdir, 'gcam_ET_para.csv'), et, delimiter=',', fmt='%.17g')
        for name, fn in (('alpha', 'gcam_albedo'), ('lai', 'gcam_lai'), ('laimin', 'gcam_laimin'),
                         ('laimax', 'gcam_laimax')):
            np.savetxt(os.path.join(pdir, fn + '.csv'), pm[name], delimiter=',', fmt='%.17g')
        lines += ['[[penman-monteith]]', 'pet_dir = pm', 'pm_tas = tas.npy', 'pm_tmin = tmin.npy', 'pm_rhs = rhs.npy',
                  'pm_rlds = rlds.npy', 'pm_rsds = rsds.npy', 'pm_wind = wind.npy', 'pm_lct = lct.npy',
                  'pm_nlcs = {}'.format(pm['nlcs']), 'pm_water_idx = {}'.format(pm['water_idx']),
                  'pm_snow_idx = {}'.format(pm['snow_idx']),
                  'pm_lc_years = ' + ', '.join(str(y) for y in pm['lc_years'])]
    elif pet == 'hs':
        hs = hs_inputs(world, start_yr, end_yr, seed=seed)
        data.update(hs)
        for k in ('hs_tas', 'hs_tmin', 'hs_tmax'):
            np.save(os.path.join(pdir, k + '.npy'), hs[k])
        lines += ['[[hargreaves-samani]]', 'pet_dir = hs', 'hs_tas = hs_tas.npy', 'hs_tmin = hs_tmin.npy',
                  'hs_tmax = hs_tmax.npy']
    elif pet == 'thornthwaite':
        tw = thornthwaite_inputs(world, start_yr, end_yr, seed=seed)
        data['trn_tas'] = tw['tair']
        np.save(os.path.join(pdir, 'tas.npy'), tw['tair'])
        lines += ['[[thornthwaite]]', 'pet_dir = thornthwaite', 'trn_tas = tas.npy']
    elif pet == 'hargreaves':
        # the step-wise v1 configuration of the reference's own integration test
        # (xanthos/test/configs/hargreaves_gwam_mrtm.ini): Hargreaves PET + GWAM runoff (+ MRTM)
        sw = stepwise_inputs(world, start_yr, end_yr, seed=seed)
        data.update(sw)
        np.save(os.path.join(pdir, 'tas.npy'), sw['temp'])
        np.save(os.path.join(pdir, 'dtr.npy'), sw['dtr'])
        lines += ['[[hargreaves]]', 'pet_dir = hargreaves', 'TemperatureFile = tas.npy',
                  'DailyTemperatureRangeFile = dtr.npy']
        gdir = os.path.join(dirs['runoff'], 'gwam')
        os.makedirs(gdir, exist_ok=True)
        lakes = np.nonzero(sw['soil_moisture'] == 999.0)[0]
        base = np.where(sw['soil_moisture'] == 999.0, 123.0, sw['soil_moisture'])     # lakes come from the two tables
        with open(os.path.join(gdir, 'max_soil_moisture.csv'), 'w') as f:
            f.write('max_soil_moisture\n')
            np.savetxt(f, base, fmt='%.17g')
        half = len(lakes) // 2
        np.savetxt(os.path.join(gdir, 'lakes.csv'), np.c_[lakes[:half] + 1, np.full(half, 999)], delimiter=',', fmt='%d')
        np.savetxt(os.path.join(gdir, 'addit_water.csv'), np.c_[lakes[half:] + 1, np.full(len(lakes) - half, 999)],
                   delimiter=',', fmt='%d')
        np.save(os.path.join(gdir, 'precip.npy'), sw['precip'])
        lines += ['[Runoff]', 'runoff_module = gwam', '[[gwam]]', 'runoff_dir = gwam',
                  'runoff_spinup = {}'.format(m if runoff_spinup is None else runoff_spinup),
                  'max_soil_moisture = max_soil_moisture.csv', 'lakes_msm = lakes.csv',
                  'addit_water_msm = addit_water.csv', 'PrecipitationFile = precip.npy']

    ab = abcd_inputs(world, m, seed=seed, with_pet=False)
    if pet != 'hargreaves':
        data.update(precip=ab['precip'], tmin=ab['tmin'], abcd_pars=ab['pars'])
    rdir = os.path.join(dirs['runoff'], 'abcd')
    np.save(os.path.join(rdir, 'pars.npy'), ab['pars'])
    np.save(os.path.join(rdir, 'precip.npy'), ab['precip'])
    np.save(os.path.join(rdir, 'tmin.npy'), ab['tmin'])
    if pet != 'hargreaves':
        lines += ['[Runoff]', 'runoff_module = abcd', '[[abcd]]', 'runoff_dir = abcd', 'calib_file = pars.npy',
                  'runoff_spinup = {}'.format(m if runoff_spinup is None else runoff_spinup), 'jobs = -1',
                  'PrecipitationFile = ' + os.path.join(rdir, 'precip.npy'),
                  'TempMinFile = ' + os.path.join(rdir, 'tmin.npy')]
    if routing:
        mdir = os.path.join(dirs['routing'], 'mrtm')
        np.save(os.path.join(mdir, 'velocity.npy'), world.velocity)
        np.save(os.path.join(mdir, 'fdistance.npy'), world.flow_dist)
        np.save(os.path.join(mdir, 'fdirection.npy'), world.flow_dir)
        lines += ['[Routing]', 'routing_module = mrtm', '[[mrtm]]', 'routing_dir = mrtm',
                  'routing_spinup = {}'.format(m if routing_spinup is None else routing_spinup),
                  'channel_velocity = velocity.npy', 'flow_distance = fdistance.npy', 'flow_direction = fdirection.npy']
    project_lines = [
        '[Project]', 'ProjectName = ' + project, 'RootDir = ' + root, 'InputFolder = input', 'OutputFolder = output',
        'RefDir = reference', 'pet_dir = pet', 'RoutingDir = routing', 'RunoffDir = runoff',
        'ncell = {}'.format(n), 'ngridrow = {}'.format(world.nrow), 'ngridcol = {}'.format(world.ncol),
        'HistFlag = True', 'n_basins = {}'.format(world.n_basins), 'StartYear = {}'.format(start_yr),
        'EndYear = {}'.format(end_yr), 'output_vars = ' + output_vars, 'OutputFormat = 4', 'OutputUnit = 0',
        'OutputInYear = 0', 'AggregateRunoffBasin = 1', 'AggregateRunoffCountry = 0', 'AggregateRunoffGCAMRegion = 0',
        'PerformDiagnostics = 0', 'CreateTimeSeriesPlot = 0',
        'CalculateDroughtStats = {}'.format(int('drought' in (postproc or {}))),
        'CalculateAccessibleWater = {}'.format(int('accessible_water' in (postproc or {}))),
        'CalculateHydropowerPotential = 0', 'CalculateHydropowerActual = 0',
        'Calibrate = {}'.format(int(bool(calibrate)))]
    if postproc and 'drought' in postproc:
        lines += ['[Drought]'] + ['{} = {}'.format(k, v) for k, v in postproc['drought'].items()]
    if postproc and 'accessible_water' in postproc:
        adir = os.path.join(inp, 'accessible_water')
        os.makedirs(adir, exist_ok=True)
        rng = np.random.default_rng(seed + 77)
        data['res_capacity'] = rng.uniform(0.0, 50.0, world.n_basins)
        data['bfi'] = rng.uniform(0.1, 0.9, world.n_basins)
        np.savetxt(os.path.join(adir, 'total_reservoir_storage_capacity_BM3.csv'), data['res_capacity'], fmt='%.17g')
        with open(os.path.join(adir, 'bfi_per_basin.csv'), 'w') as f:
            f.write('basin_id,bfi_avg\n')
            for b in range(world.n_basins):
                f.write('{},{:.17g}\n'.format(b + 1, data['bfi'][b]))
        project_lines.append('AccWatDir = accessible_water')
        lines += ['[AccessibleWater]', 'ResCapacityFile = total_reservoir_storage_capacity_BM3.csv',
                  'BfiFile = bfi_per_basin.csv'] + ['{} = {}'.format(k, v) for k, v in postproc['accessible_water'].items()]
    if calibrate:
        streamflow = calibrate == 'streamflow'      # calibrate=True: runoff target; 'streamflow': set_calibrate = 1
        lines += ['[Calibrate]', 'set_calibrate = {}'.format(int(streamflow)),
                  'observed = ' + os.path.join(root, 'input', 'obs.csv'),
                  'obs_unit = ' + ('m3_per_sec' if streamflow else 'km3_per_mth'), 'calib_out_dir = ' + os.path.join(root, 'output', 'calib'),
                  'calibration_basins = 1-{}'.format(world.n_basins)]
    if extra_lines:
        lines += list(extra_lines)
    if project_overrides:
        for key, val in project_overrides.items():
            hit = [k for k, ln in enumerate(project_lines) if ln.split('=')[0].strip() == key]
            if hit:
                project_lines[hit[0]] = '{} = {}'.format(key, val)
            else:
                project_lines.append('{} = {}'.format(key, val))
    data['country_ids'], data['region_ids'] = country, region
    ini = os.path.join(root, project + '.ini')
    with open(ini, 'w') as f:
        f.write('\n'.join(project_lines + lines) + '\n')
    return ini, data


def write_observations(path, series, start_yr):
    """Observed basin runoff CSV: header + rows basin,year,month,value (docs/calibration_tutorial.md:7-8)."""
    nb, m = series.shape
    with open(path, 'w') as f:
        f.write('basin,year,month,value\n')
        for b in range(nb):
            for k in range(m):
                f.write('{},{},{},{:.17g}\n'.format(b + 1, start_yr + k // 12, k % 12 + 1, series[b, k]))
