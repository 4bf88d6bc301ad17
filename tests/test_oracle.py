"""
CPU tests: the numpy oracle against (a) the committed golden fixtures that were
produced by running the reference, (b) the reference's own known-answer test
(xanthos/test/test_thornthwaite.py) and (c) the live reference when
/root/reference is present (build container only).
"""

import numpy as np
import pytest

from oracle import pet as opet, mrtm as omrtm, abcd as oabcd
from oracle import ref_loader
from oracle.validate_against_reference import (run_oracle, oracle_calibration, build_case, run_reference, compare,
                                               run_oracle_stepwise, build_stepwise_case, run_reference_stepwise)
from util import load_golden, bitwise_equal


@pytest.mark.parametrize("name", ["case_a", "case_b"])
def test_oracle_matches_golden_bitwise(name):
    case, ref = load_golden(name)
    out = run_oracle(case)
    for k, v in out.items():
        assert bitwise_equal(v, ref[k]), k
    ed = oracle_calibration(case, ref)
    assert np.max(np.abs(ed - ref['cal_ed']) / np.abs(ref['cal_ed'])) < 1e-12


def test_stepwise_oracle_matches_golden_bitwise():
    """Hargreaves PET + GWAM runoff with its spin-up pass (tests/golden/case_c.npz, produced by the reference)."""
    case, ref = load_golden("case_c")
    out = run_oracle_stepwise(case)
    assert set(out) == set(ref)
    for k, v in out.items():
        assert bitwise_equal(v, ref[k]), k


def test_postproc_oracle_matches_golden_bitwise():
    """Drought statistics, Aggregation_Map and the accessible-water chain (tests/golden/case_d.npz, produced by the
    reference's own functions)."""
    from oracle.validate_against_reference import run_oracle_postproc
    case, ref = load_golden("case_d")
    out = run_oracle_postproc(case)
    assert set(out) == set(ref)
    for k, v in out.items():
        assert bitwise_equal(v, ref[k]), k
    assert np.isnan(ref['thr12']).any() and (ref['dur12'].max() > 12)      # the edge cases are really in the fixture


def test_upstream_fast_equals_loop():
    case, ref = load_golden("case_a")
    a = omrtm.upstream_fast(case['coords'], ref['dsid'], case['nrow'], case['ncol'])
    assert np.array_equal(a, ref['upid'])


def test_thornthwaite_known_answer():
    """xanthos/test/test_thornthwaite.py:12-37, restated for the oracle."""
    days = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    eq = opet.daylight_hours(days, np.array([0.0]))
    npole = opet.daylight_hours(days, np.array([np.pi / 2]))
    spole = opet.daylight_hours(days, np.array([-np.pi / 2]))
    assert np.all(eq == 12)
    assert np.any(npole[0] == 0.0) and np.any(npole[0] == 24.0)
    assert np.all(24 - npole == spole)
    lat = np.array([0.698132])
    tas1 = np.array([[2, 5, 6, 8, 10, 12, 15, 12, 10, 8, 6, 5]], dtype=float)
    tas2 = -np.ones((1, 12))
    want = np.array([[9.7, 22.9, 33.7, 47.6, 66.0, 78.8, 98.5, 74.4, 54.8, 40.9, 27.1, 22.3]])
    assert np.all(np.round(opet.thornthwaite_pet(tas1, lat, 1999, 1999), 1) == want)
    assert np.all(opet.thornthwaite_pet(tas2, lat, 1999, 1999) == 0)


def test_abcd_spinup_too_short_raises():
    case, _ = load_golden("case_a")
    with pytest.raises(IndexError):
        oabcd.abcd_execute(case['n_basins'], case['basin_ids'], case['abcd_pet'], case['precip'],
                           np.nan_to_num(case['tmin']), case['abcd_pars'], case['nmonths'], 24)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_oracle_matches_live_reference():
    case = build_case(nrow=20, ncol=40, ncell=180, n_basins=5, seed=77, start_yr=2003, end_yr=2005,
                      spinup=30, routing_spinup=2)
    r = run_reference(case)
    o = run_oracle(case)
    for k, (bit, rel, same_nan) in compare(r, o, verbose=False).items():
        assert bit, (k, rel)
    sc = build_stepwise_case(nrow=20, ncol=40, ncell=180, n_basins=5, seed=78, start_yr=2003, end_yr=2004, spinup=9)
    r, o = run_reference_stepwise(sc), run_oracle_stepwise(sc)
    for k in r:
        assert bitwise_equal(r[k], o[k]), k
