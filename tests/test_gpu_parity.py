"""
GPU parity tests: the CUDA path (through the ctypes C-ABI) against the golden fixtures produced
by the reference and against the numpy oracle on seeded synthetic worlds.

Tolerances: integer topology and MRTM routing are BIT-EXACT; fp64 PET / ABCD / KGE results are
within 1e-9 relative (north_star), measured with a small absolute floor where values pass
through zero.
"""

from types import SimpleNamespace

import numpy as np
import pytest

from util import load_golden, max_rel, bitwise_equal

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _settings(case):
    return SimpleNamespace(ngridrow=int(case['nrow']), ngridcol=int(case['ncol']), ncell=int(case['ncell']),
                           nmonths=int(case['nmonths']), StartYear=int(case['start_yr']), EndYear=int(case['end_yr']))


def _pm_data(case):
    keys = ['tair_load', 'TMIN_load', 'rhs_load', 'wind_load', 'rsds_load', 'rlds_load', 'tairprev_load',
            'lct_load', 'elev', 'alpha', 'lai', 'laimin', 'laimax', 'cL', 'beta', 'rslimit', 'Tminopen',
            'Tminclose', 'VPDclose', 'VPDopen', 'RBLmin', 'RBLmax', 'rc', 'emiss']
    return SimpleNamespace(**{k: case['pm_' + k] for k in keys})


@pytest.mark.parametrize("name", ["case_a", "case_b"])
def test_hs_golden(name):
    from xanthos_b200.pet import hargreaves_samani as hs
    case, ref = load_golden(name)
    data = SimpleNamespace(coords=case['coords'], hs_tas=case['hs_tas'], hs_tmax=case['hs_tmax'], hs_tmin=case['hs_tmin'])
    out = hs.execute(_settings(case), data)
    assert max_rel(out, ref['hs_pet'], floor=1e-6) < RTOL


@pytest.mark.parametrize("name", ["case_a", "case_b"])
def test_thornthwaite_golden(name):
    from xanthos_b200.pet import thornthwaite as tw
    case, ref = load_golden(name)
    out = tw.execute(case['trn_tas'].copy(), np.radians(case['lat']), case['start_yr'], case['end_yr'])
    assert max_rel(out, ref['tw_pet'], floor=1e-6) < RTOL


def test_thornthwaite_reference_unit_test():
    """xanthos/test/test_thornthwaite.py:12-37 run against the CUDA path."""
    from xanthos_b200.pet import thornthwaite
    mth_days = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    equator = thornthwaite.calc_daylight_hours(mth_days, np.array([0.0]))
    north_pole = thornthwaite.calc_daylight_hours(mth_days, np.array([np.pi / 2]))
    south_pole = thornthwaite.calc_daylight_hours(mth_days, np.array([-np.pi / 2]))
    assert np.all(equator == 12)
    assert np.any(north_pole[0] == 0.0)
    assert np.any(north_pole[0] == 24.0)
    assert np.all(24 - north_pole == south_pole)
    lat_radians = np.array([0.698132])
    tas1 = np.array([[2, 5, 6, 8, 10, 12, 15, 12, 10, 8, 6, 5]], dtype=float)
    tas2 = -np.ones((1, 12))
    pet1_correct = np.array([[9.7, 22.9, 33.7, 47.6, 66.0, 78.8, 98.5, 74.4, 54.8, 40.9, 27.1, 22.3]])
    pet1 = thornthwaite.execute(tas1, lat_radians, 1999, 1999)
    pet2 = thornthwaite.execute(tas2, lat_radians, 1999, 1999)
    assert np.all(np.round(pet1, 1) == pet1_correct)
    assert np.all(pet2 == 0)


@pytest.mark.parametrize("name", ["case_a", "case_b"])
def test_pm_golden(name):
    from xanthos_b200.pet import penman_monteith as pm
    case, ref = load_golden(name)
    out = pm.run_pmpet(_pm_data(case), case['ncell'], int(case['pm_nlcs']), case['start_yr'], case['end_yr'],
                       int(case['pm_water_idx']), int(case['pm_snow_idx']), [int(v) for v in case['pm_lc_years']])
    assert max_rel(out, ref['pm_pet'], floor=1e-6) < RTOL


@pytest.mark.parametrize("exact", ["0", "1"])
def test_pm_full_width_vs_oracle(exact, monkeypatch):
    """All 67,420 cells x 2 years (incl. a leap year): throughput kernel and exact-order kernel vs the oracle."""
    from xanthos_b200 import synthetic
    from xanthos_b200.pet import penman_monteith as pm
    from oracle import pet as opet
    monkeypatch.setenv('XANTHOS_PM_EXACT', exact)
    w = synthetic.make_world(seed=0)
    d = synthetic.pm_inputs(w, 1971, 1972, seed=1)
    for k in ('tair_load', 'TMIN_load', 'rhs_load', 'wind_load', 'rsds_load', 'rlds_load'):
        d[k] = np.nan_to_num(d[k])
    want = opet.pm_pet(d, w.ncell, d['nlcs'], 1971, 1972, d['water_idx'], d['snow_idx'], d['lc_years'])
    got = pm.run_pmpet(SimpleNamespace(**d), w.ncell, d['nlcs'], 1971, 1972, d['water_idx'], d['snow_idx'], d['lc_years'])
    assert max_rel(got, want, floor=1e-6) < (1e-12 if exact == "1" else 1e-11)


@pytest.mark.parametrize("name", ["case_a", "case_b"])
def test_abcd_golden(name, tmp_path):
    from xanthos_b200.runoff import abcd
    case, ref = load_golden(name)
    f = str(tmp_path / 'pars.npy')
    np.save(f, case['abcd_pars'])
    tmin = np.nan_to_num(case['tmin'])
    pet, aet, q, sav = abcd.abcd_execute(n_basins=case['n_basins'], basin_ids=case['basin_ids'], pet=case['abcd_pet'],
                                         precip=case['precip'], tmin=tmin, calib_file=f, n_months=case['nmonths'],
                                         spinup_steps=case['spinup'], jobs=-1)
    assert bitwise_equal(pet, case['abcd_pet'])
    assert max_rel(aet, ref['abcd_aet'], floor=1e-6) < RTOL
    assert max_rel(q, ref['abcd_q'], floor=1e-6) < RTOL
    assert max_rel(sav, ref['abcd_sav'], floor=1e-6) < RTOL
    _, _, q2, _ = abcd.abcd_execute(n_basins=case['n_basins'], basin_ids=case['basin_ids'], pet=case['abcd_pet'],
                                    precip=case['precip'], tmin=None, calib_file=f, n_months=case['nmonths'],
                                    spinup_steps=case['spinup'], jobs=1)
    assert max_rel(q2, ref['abcd_q_nosnow'], floor=1e-6) < RTOL


def test_abcd_short_spinup_raises(tmp_path):
    from xanthos_b200.runoff import abcd
    case, _ = load_golden("case_a")
    with pytest.raises(IndexError):
        abcd.abcd_execute(n_basins=case['n_basins'], basin_ids=case['basin_ids'], pet=case['abcd_pet'],
                          precip=case['precip'], tmin=None, calib_file=case['abcd_pars'], n_months=case['nmonths'],
                          spinup_steps=24, jobs=1)


def test_abcd_class_emulate():
    """ABCD(...).emulate() attributes as read by calibrate_abcd.py:153-171."""
    from xanthos_b200.runoff import abcd
    from oracle import abcd as oabcd
    case, _ = load_golden("case_a")
    idx = np.where(case['basin_ids'] == 1)[0]
    pars = np.repeat(case['abcd_pars'][:1], len(idx), axis=0)
    tmin = np.nan_to_num(case['tmin'])
    he = abcd.ABCD(pars, case['abcd_pet'][idx], case['precip'][idx], tmin[idx], np.zeros(len(idx)), case['nmonths'],
                   case['spinup'])
    he.emulate()
    aet, q, sw, _, _ = oabcd.abcd_emulate(pars, case['abcd_pet'][idx], case['precip'][idx], tmin[idx],
                                          np.zeros(len(idx)), case['nmonths'], case['spinup'])
    assert max_rel(he.rsim, q.T, floor=1e-6) < RTOL
    assert max_rel(he.actual_et, aet.T, floor=1e-6) < RTOL
    assert max_rel(he.soil_water_storage, sw.T, floor=1e-6) < RTOL


@pytest.mark.parametrize("name", ["case_a", "case_b"])
@pytest.mark.parametrize("method", ["tree", "grid"])
def test_mrtm_golden_bitwise(name, method):
    from xanthos_b200.routing import mrtm
    from xanthos_b200 import _cuda as C
    case, ref = load_golden(name)
    s = _settings(case)
    dsid = mrtm.downstream(case['coords'], case['flow_dir'], s)
    upid = mrtm.upstream(case['coords'], dsid, s)
    assert np.array_equal(dsid, ref['dsid']) and np.array_equal(upid, ref['upid'])
    um = mrtm.upstream_genmatrix(upid, 64, 16)
    chs, avg, F = mrtm.route(um, case['runoff'], case['flow_dist'], case['velocity'], case['area'], ref['ndays'],
                             case['dt'], case['routing_spinup'],
                             method=C.MRTM_TREE if method == 'tree' else C.MRTM_GRID)
    assert bitwise_equal(chs, ref['mrtm_chs'])
    assert bitwise_equal(avg, ref['mrtm_avg'])
    assert bitwise_equal(F, ref['mrtm_F'])


def test_mrtm_streamrouting_single_month():
    from xanthos_b200.routing import mrtm
    from oracle import mrtm as omrtm
    case, ref = load_golden("case_a")
    um = mrtm.upstream_genmatrix(ref['upid'])
    rows = omrtm.gather_rows(ref['upid'])
    S0 = np.abs(np.random.default_rng(0).normal(1e6, 5e5, case['ncell']))
    want = omrtm.streamrouting(case['flow_dist'], S0, np.zeros(case['ncell']), case['velocity'], case['runoff'][:, 3],
                               case['area'], 30, 10800, rows)
    got = mrtm.streamrouting(case['flow_dist'], S0, np.zeros(case['ncell']), case['velocity'], case['runoff'][:, 3],
                             case['area'], 30, 10800, um)
    for a, b in zip(got, want):
        assert bitwise_equal(a, b)


@pytest.mark.parametrize("kw,lanes", [(dict(block_threads=32, chunk_substeps=64), 0),
                                      (dict(block_threads=64, chunk_substeps=7), 12),
                                      (dict(block_threads=256, chunk_substeps=1), 20),
                                      (dict(block_threads=128, chunk_substeps=300), 9),
                                      (dict(block_threads=512, chunk_substeps=64), 14),
                                      (dict(block_threads=640, chunk_substeps=64), 10)])
def test_mrtm_cut_trees_match_oracle(kw, lanes, monkeypatch):
    """River trees much larger than a warp: exercises the cut-edge pipeline between warps."""
    if lanes:
        monkeypatch.setenv('XANTHOS_MRTM_LANES', str(lanes))
    from xanthos_b200 import synthetic
    from xanthos_b200.routing import mrtm
    from xanthos_b200 import _cuda as C
    from oracle import mrtm as omrtm
    from oracle.calendar_utils import set_month_arrays
    w = synthetic.make_world(40, 80, 1500, 8, seed=5, coast_pull=0.0)
    s = w.settings()
    m = 14
    q = synthetic.runoff_input(w, m, seed=9)
    dsid = mrtm.downstream(w.coords, w.flow_dir, s)
    upid = mrtm.upstream(w.coords, dsid, s)
    um = mrtm.upstream_genmatrix(upid, **kw)
    info = um.info
    assert info['is_forest'] == 1 and info['n_cut_edges'] > 0 and info['n_levels'] > 1
    ndays = set_month_arrays(24, 2003, 2004)[:m, 2]
    rows = omrtm.gather_rows(omrtm.upstream_fast(w.coords, omrtm.downstream(w.coords, w.flow_dir, w.nrow, w.ncol),
                                                 w.nrow, w.ncol))
    want = omrtm.route(q, w.flow_dist, w.velocity, w.area, ndays, 10800, rows, 6)
    got = mrtm.route(um, q, w.flow_dist, w.velocity, w.area, ndays, 10800, 6, method=C.MRTM_TREE)
    for a, b in zip(got, want):
        assert bitwise_equal(a, b)
    got = mrtm.route(um, q, w.flow_dist, w.velocity, w.area, ndays, 3600.0, 0, method=C.MRTM_TREE)   # hourly variant
    want = omrtm.route(q[:, :3], w.flow_dist, w.velocity, w.area, ndays, 3600.0, rows, 0)
    assert bitwise_equal(got[0][:, :3], want[0]) and bitwise_equal(got[1][:, :3], want[1])


def test_mrtm_member_batch_is_bitwise_the_single_member_run():
    """xan_mrtm_route_batch: two and three members per call (two per warp pass) == one call per member."""
    from xanthos_b200 import synthetic
    from xanthos_b200.routing import mrtm
    from xanthos_b200 import _cuda as C
    from oracle.calendar_utils import set_month_arrays
    w = synthetic.make_world(40, 80, 1500, 8, seed=5, coast_pull=0.0)
    s = w.settings()
    m = 14
    um = mrtm.upstream_genmatrix(mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s))
    assert um.info['n_cut_edges'] > 0
    ndays = set_month_arrays(24, 2003, 2004)[:m, 2]
    qs = [C.Field.from_host(synthetic.runoff_input(w, m, seed=20 + k)) for k in range(3)]
    single = [mrtm.route_device(um, q, w.flow_dist, w.velocity, w.area, ndays, 10800, 6) for q in qs]
    for k in (2, 3):
        batch = mrtm.route_device_batch(um, qs[:k], w.flow_dist, w.velocity, w.area, ndays, 10800, 6)
        for (c1, a1, i1), (c2, a2, i2) in zip(single[:k], batch):
            assert bitwise_equal(c1.to_host(), c2.to_host()) and bitwise_equal(a1.to_host(), a2.to_host())
            assert bitwise_equal(i1.cpu().numpy(), i2.cpu().numpy())


def test_hargreaves_pet_matches_golden():
    """Series launch and the reference's per-month signature (hargreaves.py:17-39) against case_c."""
    from xanthos_b200.pet import hargreaves as hg
    from xanthos_b200.utils.general import set_month_arrays
    case, ref = load_golden('case_c')
    ymd = set_month_arrays(case['nmonths'], case['start_yr'], case['end_yr'])
    lat = np.radians(case['lat'])
    pet = hg.series_device(case['temp'], case['dtr'], lat, ref['solar_dec'], ref['dr'], ymd[:, 2]).to_host()
    assert max_rel(pet, ref['pet'], floor=1e-6) < RTOL
    assert np.array_equal(pet == 0, ref['pet'] == 0)                # clipped months and NaN -> 0 inputs agree exactly
    k = 7
    dtr = np.nan_to_num(case['dtr'][:, k])
    one = hg.calculate_pet(np.nan_to_num(case['temp'][:, k]), dtr, lat, ref['solar_dec'][k], ref['dr'][k], ymd[k, 2])
    assert one.shape == (case['ncell'],) and max_rel(one, ref['pet'][:, k], floor=1e-6) < RTOL
    assert (dtr >= 0).all()                                         # mutated in place like the reference (:32)


def test_gwam_matches_golden():
    """Fused spin-up + simulation, the two-call form Components uses, and the per-month signature (gwam.py:18-88)."""
    from types import SimpleNamespace
    from xanthos_b200.runoff import gwam
    case, ref = load_golden('case_c')
    n, m, spin = case['ncell'], case['nmonths'], int(case['spinup'])
    res = gwam.run_device(ref['pet'], case['precip'], case['soil_moisture'], case['sm_prev'], m, spin)
    for k in ('aet', 'q', 'sav'):
        got = res[k].to_host()
        assert max_rel(got, ref[k], floor=1e-6) < RTOL, k
        assert np.array_equal(got == 0, ref[k] == 0), k             # branch decisions (lakes, no soil, NaN forcing)
    assert max_rel(res['sm_after_spinup'].cpu().numpy(), ref['sm_after_spinup'], floor=1e-6) < RTOL
    a = gwam.run_device(ref['pet'], case['precip'], case['soil_moisture'], case['sm_prev'], spin, 0, want=())
    b = gwam.run_device(ref['pet'], case['precip'], case['soil_moisture'], a['sm_last'].cpu().numpy(), m, 0)
    for k in ('aet', 'q', 'sav'):
        assert bitwise_equal(b[k].to_host(), res[k].to_host()), k
    rg = gwam.runoffgen(ref['pet'][:, 0], case['precip'][:, 0], SimpleNamespace(ncell=n), case['soil_moisture'],
                        ref['sm_after_spinup'])
    assert rg[0] is not None and len(rg) == 4
    for got, k in zip(rg[1:], ('aet', 'q', 'sav')):
        assert max_rel(got, ref[k][:, 0], floor=1e-6) < RTOL, k


def test_transposes_roundtrip():
    from xanthos_b200 import _cuda as C
    rng = np.random.default_rng(0)
    a = rng.normal(size=(1237, 61))
    a[5, 7] = np.nan
    a[9, 3] = np.inf
    f = C.Field.from_host(a)
    assert bitwise_equal(f.to_host(), a)
    g = C.Field.from_host(a, nan_to_num=True)
    assert bitwise_equal(g.to_host(), np.nan_to_num(a))
