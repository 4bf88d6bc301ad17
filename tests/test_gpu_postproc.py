"""
GPU tests of the post-processing scans through the C-ABI: drought thresholds and
statistics, Aggregation_Map, accessible water.  Every result is compared BITWISE with the golden outputs of the
reference's own functions (tests/golden/case_d.npz) and with the numpy oracle on larger seeded inputs.
"""

import os
from types import SimpleNamespace

import numpy as np
import pytest

from util import load_golden, bitwise_equal

pytestmark = pytest.mark.gpu


def _settings(case, nper, **kw):
    return SimpleNamespace(threshold_start_year=int(case['thr_start']), threshold_end_year=int(case['thr_end']),
                           StartYear=int(case['start_yr']), EndYear=int(case['end_yr']), threshold_nper=nper, **kw)


@pytest.mark.parametrize("nper", [1, 12])
def test_drought_thresholds_and_stats_match_reference_bitwise(nper):
    from xanthos_b200.drought.drought_stats import DroughtStats
    case, ref = load_golden("case_d")
    h = np.ascontiguousarray(case['hydro'].T)                         # [ntime, ngrid], as the reference's methods take it
    thr = DroughtStats.calculate_thresholds(h, _settings(case, nper))
    assert bitwise_equal(thr, ref['thr%d' % nper])
    S, I, D = DroughtStats.droughtstats(None, h, thr)
    for got, key in ((S, 'sev'), (I, 'int'), (D, 'dur')):
        assert bitwise_equal(got, ref['%s%d' % (key, nper)]), key
    # getthresh on its own, 11 samples -> the virtual index (n - 1) q is the integer 1
    from oracle import postproc as P
    assert bitwise_equal(DroughtStats.getthresh(h[:11], 1), P.getthresh(h[:11], 1))
    with pytest.raises(ValueError):
        DroughtStats.getthresh(h[:13], 12)
    if nper == 1:      # one period over the whole record: 372 samples per cell (the local-memory path of the kernel)
        assert bitwise_equal(DroughtStats.getthresh(h, 1), P.getthresh(h, 1))


def test_drought_module_from_resident_field_writes_reference_files(tmp_path):
    """DroughtStats(settings, Q, Sav) as Components.drought calls it: thresholds .npy first, then the statistics files."""
    from xanthos_b200.drought.drought_stats import DroughtStats
    from xanthos_b200 import _cuda as C
    case, ref = load_golden("case_d")
    q = np.ascontiguousarray(case['hydro'])                           # [ncell, nmonths]
    q = C.remember(q, C.Field.from_host(q))                           # as if a CUDA stage had returned it
    s = _settings(case, 12, drought_var='q', drought_thresholds=None, OutputFolder=str(tmp_path), OutputNameStr='t',
                  output_vars=[], ProjectName='t', OutputFormat=4, OutputUnit=0, OutputUnitStr='mmpermonth',
                  OutputInYear=0)
    DroughtStats(s, q, None)
    thr_file = os.path.join(str(tmp_path), 'drought_thresholds_t.npy')
    assert bitwise_equal(np.load(thr_file), ref['thr12'])
    s.drought_thresholds = thr_file
    DroughtStats(s, q, None)
    for name, key in (('severity', 'sev12'), ('intensity', 'int12'), ('duration', 'dur12')):
        got = np.load(os.path.join(str(tmp_path), 'drought_{}_t.npy'.format(name)))
        assert bitwise_equal(got, ref[key].T), name                  # written [ngrid, ntime] like the reference
    s.drought_var = 'pet'
    with pytest.raises(ValueError):
        DroughtStats(s, q, None)


def test_aggregation_map_matches_reference_bitwise():
    from xanthos_b200.diagnostics.time_series import Aggregation_Map
    case, ref = load_golden("case_d")
    assert bitwise_equal(Aggregation_Map(case['ids'], case['hydro']), ref['aggmap'])


def test_accessible_water_matches_reference_bitwise(tmp_path):
    import pandas as pd
    from xanthos_b200.accessible import accessible as acc
    case, ref = load_golden("case_d")
    nb = int(case['ids'].max())
    mr = acc.basin_annual_runoff_device(case['hydro'], case['area'], case['ids']).cpu().numpy()
    assert bitwise_equal(mr, ref['basin_annual'])
    pd.DataFrame(case['res_capacity']).to_csv(tmp_path / 'res.csv', header=False, index=False)
    pd.DataFrame({'bfi_avg': case['bfi']}).to_csv(tmp_path / 'bfi.csv', index=False)
    s = SimpleNamespace(ResCapacityFile=str(tmp_path / 'res.csv'), BfiFile=str(tmp_path / 'bfi.csv'),
                        nmonths=int(case['nmonths']), ncell=int(case['ncell']), MovingMeanWindow=int(case['window']),
                        StartYear=int(case['start_yr']), EndYear=int(case['end_yr']), HistEndYear=int(case['hist_end']),
                        GCAM_StartYear=int(case['gcam_start']), GCAM_EndYear=int(case['gcam_end']),
                        GCAM_YearStep=int(case['gcam_step']), Env_FlowPercent=float(case['env_flow']),
                        OutputFolder=str(tmp_path), OutputNameStr='t')
    data = SimpleNamespace(basin_names=np.array(['b%d' % i for i in range(nb)]), area=case['area'], basin_ids=case['ids'])
    # the .csv tables round-trip through text: compare the arithmetic on the parsed values
    res = pd.read_csv(s.ResCapacityFile, header=None).values
    bfi = pd.read_csv(s.BfiFile)['bfi_avg'].values
    ac = acc.AccessibleWater(s, data, case['hydro'])
    if np.array_equal(res, case['res_capacity']) and np.array_equal(bfi, case['bfi']):
        assert bitwise_equal(ac, ref['accessible'])
    else:
        assert np.allclose(ac, ref['accessible'], rtol=1e-12, atol=0)
    out = pd.read_csv(tmp_path / 'accessible_water_km3peryr_t.csv')
    assert list(out.columns[:2]) == ['id', 'name'] and out.shape == (nb, 2 + ac.shape[1])
    assert np.allclose(out.iloc[:, 2:].values, ac, rtol=1e-15)


def test_postproc_full_size_against_oracle():
    """67,420 cells x 360 months (BASELINE.json shape): bitwise against the numpy oracle, plus properties."""
    from xanthos_b200.drought.drought_stats import DroughtStats
    from xanthos_b200.diagnostics.time_series import Aggregation_Map
    from oracle import postproc as P
    rng = np.random.default_rng(5)
    n, m = 67420, 360
    h = np.abs(rng.normal(50.0, 40.0, (m, n)))                        # [ntime, ngrid]
    h[:, ::977] = 0.0
    h[17, 5::1301] = np.nan
    thr = DroughtStats.getthresh(h, 12)
    assert bitwise_equal(thr, P.getthresh(h, 12))
    S, I, D = DroughtStats.droughtstats(None, h, thr)
    So, Io, Do = P.droughtstats(h, thr)
    assert bitwise_equal(S, So) and bitwise_equal(I, Io) and bitwise_equal(D, Do)
    wet = ~(h < thr[np.arange(m) % 12])
    assert (D[wet] == 0).all() and (S[wet] == 0).all()                # not under drought -> all three are zero
    assert (np.diff(D, axis=0)[~wet[1:]] == 1).all()                  # duration counts consecutive months
    ids = rng.integers(0, 236, n)
    got = Aggregation_Map(ids, np.ascontiguousarray(h.T))
    assert bitwise_equal(got, P.aggregation_map(ids, h.T))
    assert np.allclose(got.sum(axis=0), np.nansum(h.T[ids > 0], axis=0), rtol=1e-12)


def test_run_model_with_drought_and_accessible_water(tmp_path):
    """Xanthos(ini).execute() with the two post-processing modules switched on: the files they write equal the oracle
    applied to the run's own runoff (which is checked against the oracle pipeline elsewhere)."""
    import pandas as pd
    import xanthos_b200
    from xanthos_b200 import synthetic
    from oracle import postproc as P
    w = synthetic.make_world(24, 48, 320, 6, seed=22)
    sy, ey = 2001, 2012
    post = {'drought': {'drought_var': 'q', 'threshold_nper': 12, 'threshold_start_year': 2002, 'threshold_end_year': 2009},
            'accessible_water': {'HistEndYear': 2006, 'GCAM_StartYear': 2002, 'GCAM_EndYear': 2012, 'GCAM_YearStep': 5,
                                 'MovingMeanWindow': 5, 'Env_FlowPercent': 0.1}}
    ini, data = synthetic.write_example(str(tmp_path), w, sy, ey, pet='hs', routing=False, runoff_spinup=30, postproc=post)
    res = xanthos_b200.Xanthos(ini).execute()
    out = os.path.join(str(tmp_path), 'output', 'synthetic')
    thr = np.load(os.path.join(out, 'drought_thresholds_synthetic.npy'))
    assert bitwise_equal(thr, P.calculate_thresholds(res.Q.T, sy, 2002, 2009, 12))
    ac = pd.read_csv(os.path.join(out, 'accessible_water_km3peryr_synthetic.csv'))
    mr = P.aggregation_map(w.basin_ids, P.yearly_km3(res.Q, w.area))
    want = P.accessible_water_chain(mr, sy, ey, 2006, [2002, 2007, 2012], 5, data['bfi'], data['res_capacity'], 0.1)
    assert list(ac['id']) == list(range(1, w.n_basins + 1)) and list(ac.columns[2:]) == ['2002', '2007', '2012']
    assert np.allclose(ac.iloc[:, 2:].values, want, rtol=1e-12, atol=0)
    # the basin aggregate OutWriter writes (AggregateRunoffBasin = 1) comes from the resident runoff (xan_group_sum)
    agg = np.loadtxt(os.path.join(out, 'Basin_runoff_mmpermonth_synthetic.csv'), delimiter=',', skiprows=1)
    assert np.array_equal(agg[:, 0], np.arange(1, w.n_basins + 1))
    assert bitwise_equal(agg[:, 1:], P.aggregation_map(w.basin_ids, res.Q))
