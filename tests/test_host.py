"""
CPU tests of the host-side logic: ini parser, loader, C-ABI symbols, integer topology (host code of the
library), packing invariants, sharding, the batched differential evolution driver, repository layout rules.
"""

import ctypes
import os
import re
import subprocess
import sys
from types import SimpleNamespace

import numpy as np
import pytest

from util import load_golden, bitwise_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from xanthos_b200 import _cuda as C
    hdr = open(os.path.join(ROOT, 'include', 'xanthos_b200.h')).read()
    declared = set(re.findall(r'\b(xan_[a-z0-9_]+)\s*\(', hdr))
    declared -= {'xan_pm_tables', 'xan_abcd_plan', 'xan_mrtm_plan'}
    assert declared, "no declarations found"
    lib = ctypes.CDLL(C.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(C.SIGNATURES), (declared ^ set(C.SIGNATURES))
    assert lib.xan_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from xanthos_b200.pet import hargreaves_samani as hs
    from types import SimpleNamespace
    case, _ = load_golden('case_a')
    data = SimpleNamespace(coords=case['coords'], hs_tas=case['hs_tas'], hs_tmax=case['hs_tmax'], hs_tmin=case['hs_tmin'])
    with pytest.raises(RuntimeError):
        hs.execute(SimpleNamespace(StartYear=1999, EndYear=2001), data)


def test_product_never_imports_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, 'xanthos_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(d, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


@pytest.mark.parametrize("name", ["case_a", "case_b"])
def test_host_topology_bit_exact(name):
    """downstream / upstream / UM are integer host code of the library: checked here without a GPU."""
    from types import SimpleNamespace
    from xanthos_b200.routing import mrtm
    case, ref = load_golden(name)
    s = SimpleNamespace(ngridrow=int(case['nrow']), ngridcol=int(case['ncol']))
    dsid = mrtm.downstream(case['coords'], case['flow_dir'], s)
    upid = mrtm.upstream(case['coords'], dsid, s)
    assert np.array_equal(dsid, ref['dsid'])
    assert np.array_equal(upid, ref['upid'])
    um = mrtm.upstream_genmatrix(upid)
    indptr, indices, data = um.csr_arrays()
    assert np.array_equal(indptr, ref['um_indptr'])
    assert np.array_equal(indices, ref['um_indices'])
    assert np.array_equal(data, ref['um_data'])


@pytest.mark.parametrize("lanes", [0, 9, 16])
def test_mrtm_packing_invariants(lanes, monkeypatch):
    from xanthos_b200 import synthetic
    from xanthos_b200.routing import mrtm
    if lanes:
        monkeypatch.setenv('XANTHOS_MRTM_LANES', str(lanes))
    w = synthetic.make_world(40, 80, 1500, 8, seed=5, coast_pull=0.0)
    s = w.settings()
    dsid = mrtm.downstream(w.coords, w.flow_dir, s)
    upid = mrtm.upstream(w.coords, dsid, s)
    um = mrtm.upstream_genmatrix(upid)
    info = um.info
    lane_cell, ep, ec = um.packing()
    assert info['is_forest'] == 1
    assert np.array_equal(np.sort(lane_cell[lane_cell >= 0]), np.arange(w.ncell))      # every cell exactly once
    assert (ep < ec).all()                                                             # producers precede consumers
    assert len(ep) == info['n_cut_edges']
    # a warp has at most one consumer warp (no back-pressure cycles, see csrc/mrtm.cu)
    cons = {}
    for p, c in zip(ep, ec):
        cons.setdefault(int(p), set()).add(int(c))
    assert all(len(v) == 1 for v in cons.values())
    # cells + ghost lanes fit a warp
    ghosts = np.bincount(ec, minlength=lane_cell.shape[0]) if len(ec) else np.zeros(lane_cell.shape[0], dtype=int)
    assert ((lane_cell >= 0).sum(axis=1) + ghosts <= 32).all()
    # every cut edge joins a cell to its receiver
    warp_of = np.empty(w.ncell, dtype=int)
    for wi in range(lane_cell.shape[0]):
        warp_of[lane_cell[wi][lane_cell[wi] >= 0]] = wi
    has = dsid > 0
    up_is_adjacent = np.zeros(w.ncell, dtype=bool)
    for i in range(w.ncell):
        for j in upid[i, :upid[i, 8]]:
            up_is_adjacent[j - 1] = True
    crossing = [(warp_of[i], warp_of[dsid[i] - 1]) for i in range(w.ncell)
                if has[i] and up_is_adjacent[i] and warp_of[i] != warp_of[dsid[i] - 1]]
    assert sorted(crossing) == sorted(zip(ep.tolist(), ec.tolist()))


def test_cyclic_graph_is_not_a_forest():
    from xanthos_b200.routing import mrtm
    upid = np.zeros((3, 9), dtype=np.int64)
    upid[0, 0], upid[0, 8] = 2, 1      # 2 -> 1
    upid[1, 0], upid[1, 8] = 3, 1      # 3 -> 2
    upid[2, 0], upid[2, 8] = 1, 1      # 1 -> 3 : a cycle
    info = mrtm.upstream_genmatrix(upid).info
    assert info['is_forest'] == 0 and info['n_warps'] == 0


def test_ini_parser_and_config_reader(tmp_path):
    from xanthos_b200 import synthetic
    from xanthos_b200.data_reader.ini_reader import ConfigReader, parse_ini
    from xanthos_b200.data_reader.data_load import DataLoader, ValidationException
    w = synthetic.make_world(18, 36, 150, 4, seed=2)
    ini, data = synthetic.write_example(str(tmp_path), w, 1999, 2001, pet='pm', routing_spinup=5)
    raw = parse_ini(ini)
    assert raw['PET']['penman-monteith']['pm_lc_years'] == ['1970', '1975', '1980', '1985', '1990', '1995', '2000']
    c = ConfigReader(ini)
    assert (c.ncell, c.ngridrow, c.ngridcol, c.nmonths) == (150, 18, 36, 36)
    assert c.mod_cfg == 'pm_abcd_mrtm' and c.routing_spinup == 5 and c.pm_lc_years[-1] == 2000
    assert c.output_vars == ['q', 'avgchflow'] and c.OutputUnitStr == 'mmpermonth'
    d = DataLoader(c)
    assert d.tair_load.shape == (150, 36) and d.basin_ids.dtype.kind == 'i'
    assert np.allclose(d.area, w.area) and np.array_equal(d.flow_dir, w.flow_dir)
    assert np.array_equal(d.lct_load, data['lct_load'])
    c.update({'pm_tas': np.zeros((10, 36))})
    with pytest.raises(ValidationException):
        DataLoader(c)


def test_set_month_arrays_mod4_rule():
    from xanthos_b200.utils.general import set_month_arrays
    from oracle.calendar_utils import set_month_arrays as oracle_sma
    a = set_month_arrays(36, 2098, 2100)
    assert np.array_equal(a, oracle_sma(36, 2098, 2100))
    assert a[25, 2] == 29        # February 2100 is a leap month under the mod-4 rule


def test_sinusoidal_factor_bitwise():
    """utils/general.py::calc_sinusoidal_factor (host logic of the Hargreaves PET) against the reference's output."""
    from xanthos_b200.utils.general import calc_sinusoidal_factor, set_month_arrays
    case, ref = load_golden('case_c')
    sd, dr = calc_sinusoidal_factor(set_month_arrays(case['nmonths'], case['start_yr'], case['end_yr']))
    assert np.array_equal(sd, ref['solar_dec']) and np.array_equal(dr, ref['dr'])


def test_stepwise_project_config(tmp_path):
    """hargreaves + gwam + mrtm project (the configuration of xanthos/test/test_hargreaves_gwam_mrtm.py) from disk."""
    from xanthos_b200 import synthetic
    from xanthos_b200.data_reader.ini_reader import ConfigReader
    from xanthos_b200.data_reader.data_load import DataLoader
    from xanthos_b200.configurations import ConfigRunner
    w = synthetic.make_world(24, 48, 320, 6, seed=21)
    ini, data = synthetic.write_example(str(tmp_path), w, 2003, 2005, pet='hargreaves', routing_spinup=7, runoff_spinup=14)
    c = ConfigReader(ini)
    assert (c.pet_module, c.runoff_module, c.routing_module, c.runoff_spinup) == ('hargreaves', 'gwam', 'mrtm', 14)
    d = DataLoader(c)
    assert np.array_equal(d.soil_moisture, data['soil_moisture']) and np.array_equal(d.sm_prev, 0.5 * data['soil_moisture'])
    assert np.array_equal(d.precip, data['precip'], equal_nan=True)
    assert (d.dtr[~np.isnan(d.dtr)] >= 0).all()                     # neg_to_zero (data_load.py:83-84)
    r = ConfigRunner(c)
    assert r.spinup and r.pet_timestep == 36 and r.runoff_timestep == 36 and r.routing_timestep == 0


def test_sharding_partitions():
    from xanthos_b200 import sharding, synthetic
    from xanthos_b200.routing import mrtm
    w = synthetic.make_world(36, 72, 900, 12, seed=3)
    parts = sharding.partition_basins(w.basin_ids, 4)
    allb = np.sort(np.concatenate(parts))
    assert np.array_equal(allb, np.arange(1, w.n_basins + 1))
    counts = np.bincount(w.basin_ids)
    loads = [counts[p].sum() for p in parts]
    assert max(loads) - min(loads) <= counts.max()
    assert [len(p) for p in sharding.partition_members(64, 8)] == [8] * 8
    assert sum(len(p) for p in sharding.partition_members(10, 4)) == 10
    dsid = mrtm.downstream(w.coords, w.flow_dir, w.settings())
    assert sharding.basins_closed_under_flow(dsid, w.basin_ids)
    wc = synthetic.make_world(36, 72, 900, 12, seed=3, cut_basins=True)
    assert not sharding.basins_closed_under_flow(mrtm.downstream(wc.coords, wc.flow_dir, wc.settings()), wc.basin_ids)
    sh = sharding.BasinShard(w.basin_ids, parts[1])
    assert np.all(np.isin(w.basin_ids[sh.cells], parts[1]))
    tair = np.arange(w.ncell, dtype=float)[:, None] * np.ones((1, 3))
    ext = sh.take_with_halo(tair)
    want_prev = np.where(sh.cells > 0, tair[np.maximum(sh.cells - 1, 0), 0], 0.0)
    got_prev = np.where(sh.prev_idx >= 0, ext[np.maximum(sh.prev_idx, 0), 0], 0.0)
    assert np.array_equal(got_prev, want_prev)


def test_batched_differential_evolution_converges():
    from xanthos_b200.calibrate.calibrate_abcd import differential_evolution_batched, BOUNDS_SNOW, expand_str_range
    target = np.array([[0.3, 2.0, 0.5, 0.7, 0.1], [0.9, 7.0, 0.2, 0.1, 0.6]])
    calls = []

    def ev(x, idx):
        calls.append(len(idx))
        return np.sqrt(((x - target[idx][:, None, :]) ** 2).sum(axis=2))
    r = differential_evolution_batched(ev, 2, BOUNDS_SNOW, seed=1, maxiter=400)
    assert np.abs(r['x'] - target).max() < 1e-3
    assert expand_str_range(['0-2', '6', '7-9']) == [0, 1, 2, 6, 7, 8, 9]


GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from xanthos_b200 import sharding
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%s' % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
basin_ids = np.repeat(np.arange(1, 7), [5, 9, 2, 7, 4, 3])
parts = sharding.partition_basins(basin_ids, 2)
mine = parts[r]
vals = torch.tensor([[float(b), float(b) * 10] for b in mine], dtype=torch.float64)
full = sharding.gather_ragged_rows(mine - 1, vals, 6)
assert torch.equal(full[:, 0], torch.arange(1, 7, dtype=torch.float64)), full
agg = torch.full((3, 4), float(r), dtype=torch.float64)
st = sharding.gather_stack(agg)
assert st.shape == (2, 3, 4) and float(st[0].mean()) == 0.0 and float(st[1].mean()) == 1.0
members = sharding.partition_members(5, 2)[r]
assert list(members) == ([0, 1, 2] if r == 0 else [3, 4])
# calibrate_all under torch.distributed: whole basins per rank, result files per basin, tables gathered on every rank
# (the differential evolution itself needs the GPU: it is replaced by a deterministic stand-in here)
from types import SimpleNamespace
from xanthos_b200.calibrate import calibrate_abcd as cal
seen = []
def fake(basin_nums, *a, **k):
    seen.extend(basin_nums)
    b = np.asarray(basin_nums, dtype=float)
    return np.stack([b, b * 2, b * 3, b * 4, b * 5], axis=1), b / 10.0, {'nfev': np.full(len(b), 7)}
cal.calibrate_basins = fake
st = SimpleNamespace(set_calibrate=0, cal_basins=['1-4', '6'], nmonths=12, runoff_spinup=12, obs_unit='km3_per_mth',
                     calib_out_dir=sys.argv[4])
data = SimpleNamespace(basin_ids=basin_ids, basin_names=None, area=None, precip=None, cal_obs=None, tmin=np.zeros(1))
pars, kge = cal.calibrate_all(st, data, None, None)
want = np.array([1., 2., 3., 4., 6.])
assert pars.shape == (5, 5) and np.array_equal(pars[:, 0], want) and np.allclose(kge, want / 10), (pars, kge)
assert sorted(seen) == sorted(int(b) for b in sharding.partition_basins(basin_ids, 2, basins=[1, 2, 3, 4, 6])[r])
dist.barrier()
for b in (1, 2, 3, 4, 6):
    assert np.load(os.path.join(sys.argv[4], 'kge_result_basin_%d.npy' % b))[0] == b / 10.0
    assert np.array_equal(np.load(os.path.join(sys.argv[4], 'abcdm_parameters_basin_%d.npy' % b))[0, :2], [b, 2 * b])
assert not os.path.exists(os.path.join(sys.argv[4], 'kge_result_basin_5.npy'))
dist.destroy_process_group()
print("rank", r, "ok")
'''


def test_two_rank_gather_with_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all('ok' in o for o in outs), outs


@pytest.mark.parametrize("fmt,ext", [(0, 'nc'), (1, 'csv'), (2, 'mat'), (3, 'parquet'), (4, 'npy')])
def test_out_writer_formats_round_trip(tmp_path, fmt, ext):
    """OutWriter file formats (out_writer.py:159-235) and the matching DataLoader.load_data readers (data_load.py:343-390);
    host-only code, no device needed for arrays that are not resident."""
    from types import SimpleNamespace
    from xanthos_b200.data_writer.out_writer import OutWriter
    from xanthos_b200.data_reader.data_load import DataLoader
    a = np.random.default_rng(fmt).random((40, 24)) * 100
    s = SimpleNamespace(output_vars=['q', 'avgchflow'], ProjectName='p', OutputFolder=str(tmp_path), OutputFormat=fmt,
                        OutputUnit=1, OutputUnitStr='km3permonth', OutputInYear=1, StartYear=2001, EndYear=2002)
    area = np.linspace(1000.0, 3000.0, 40)
    OutWriter(s, area, {'q': a.copy(), 'avgchflow': a.copy()}).write()
    want_q = a.reshape(40, 2, 12).sum(axis=2) * (area / 1e6)[:, None]          # yearly sum, mm -> km3
    want_f = a.reshape(40, 2, 12).mean(axis=2)                                  # streamflow: yearly mean, no conversion
    fq = os.path.join(str(tmp_path), 'q_km3permonth_p.' + ext)
    ff = os.path.join(str(tmp_path), 'avgchflow_m3persec_p.' + ext)
    assert os.path.isfile(fq) and os.path.isfile(ff)
    if ext == 'nc':
        got_q, got_f = DataLoader.load_data(fq, key='data'), DataLoader.load_data(ff, key='data')
        assert got_q.dtype == np.float32 and np.allclose(got_q, want_q, rtol=1e-6) and np.allclose(got_f, want_f, rtol=1e-6)
    elif ext == 'mat':
        assert np.array_equal(DataLoader.load_data(fq, key='q'), want_q)
        assert np.array_equal(DataLoader.load_data(ff, key='avgchflow'), want_f)
    elif ext == 'npy':
        assert np.array_equal(DataLoader.load_data(fq), want_q) and np.array_equal(DataLoader.load_data(ff), want_f)
    elif ext == 'csv':
        got = np.loadtxt(fq, delimiter=',', skiprows=1)
        assert np.array_equal(got[:, 0], np.arange(1, 41)) and np.array_equal(got[:, 1:], want_q)
        assert open(fq).readline().strip() == 'id,2001,2002'
    else:
        import pyarrow.parquet as pq
        t = pq.read_table(fq).to_pandas()
        assert list(t.columns) == ['id', '2001', '2002'] and np.array_equal(t.iloc[:, 1:].values, want_q)


def test_skew_plan_and_schedule_model_equal_the_oracle_bitwise(monkeypatch):
    """The plan of the skew routing kernel (csrc/mrtm_skew.cu: pieces, places, per-cell lags, ghost / export entries)
    driven through the CPU model of the kernel's schedule (tests/skew_model.py) reproduces oracle.mrtm.route bit for
    bit, for 2 and 4 cells per lane; structural invariants of the tables."""
    import skew_model
    from xanthos_b200 import synthetic
    from xanthos_b200.routing import mrtm
    from oracle import mrtm as omrtm
    from oracle.calendar_utils import set_month_arrays
    w = synthetic.make_world(30, 60, 900, 4, seed=7)
    s = w.settings()
    up = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
    months, spin = 4, 2
    q = synthetic.runoff_input(w, months, seed=3)
    nd = set_month_arrays(12, 1971, 1971)[:months, 2]
    oup = omrtm.upstream_fast(w.coords, omrtm.downstream(w.coords, w.flow_dir, w.nrow, w.ncol), w.nrow, w.ncol)
    for K in ('1', '2', '4'):
        monkeypatch.setenv('XANTHOS_MRTM_SKEW_K', K)
        um = mrtm.upstream_genmatrix(up)
        t = skew_model.skew_tables(um)
        inf = t['info']
        assert inf['K'] == int(K) and inf['n_warps'] > 0
        cells = t['cell'][t['cell'] >= 0]
        assert np.array_equal(np.sort(cells), np.arange(w.ncell))                 # every cell has exactly one place
        assert (t['lag'] >= 0).all() and t['lag'].max() <= inf['Dmax'] and (t['src'] <= inf['zero']).all()
        wide = np.bincount(up[:, 8].astype(int) >= 2, minlength=2)[1]
        assert ((t['cell'][:, :32] >= 0).sum() >= wide)                            # rows with >= 2 tributaries sit in slot 0
        for dt in (10800.0, 21600.0):                                             # 21600 s: many cells empty at every other sub-step
            got = skew_model.route(t, q, w.flow_dist, w.velocity, w.area, nd, dt, spin)
            want = omrtm.route(q, w.flow_dist, w.velocity, w.area, nd, dt, omrtm.csr_rows(oup), spin)
            for a, b, name in zip(got, want, ('ChStorage', 'Avg_ChFlow', 'instream_flow')):
                assert bitwise_equal(a, b), (K, dt, name)


def test_outlet_cells_of_the_streamflow_target_match_the_oracle():
    """Host logic of the streamflow calibration target (INTENDED semantics, oracle/calibrate.py): the outlet of a basin
    is its cell with the largest drainage area - vectorised product code against the sequential oracle."""
    from xanthos_b200 import synthetic
    from xanthos_b200.calibrate import calibrate_abcd as cal
    from oracle import calibrate as ocal, mrtm as omrtm
    for args in [(24, 48, 320, 5, 44), (40, 80, 1500, 8, 5), (90, 180, 6000, 30, 2)]:
        w = synthetic.make_world(*args[:4], seed=args[4])
        dsid = omrtm.downstream(w.coords, w.flow_dir, w.nrow, w.ncol)
        a, b = cal.outlet_cells(w.basin_ids, dsid, w.area), ocal.outlet_cells(w.basin_ids, dsid, w.area)
        assert np.array_equal(a, b)
        assert (w.basin_ids[a] == np.arange(1, w.n_basins + 1)).all()          # the outlet of basin b lies in basin b
        down = dsid[a] - 1                                                      # and drains out of the basin (or nowhere)
        assert all(d < 0 or w.basin_ids[d] != w.basin_ids[c] for c, d in zip(a, down))


def test_routing_raster_vectorize_matches_the_oracle_and_the_live_reference():
    """f2: `DataLoader.load_routing_data` for a 2-D DRT raster (flip, 68-row offset, Fortran-order sampling, rep_val;
    data_load.py:392-425) against the oracle restatement, and the oracle against the reference's own `vectorize` /
    `sub2ind` where the reference tree is present."""
    from xanthos_b200.data_reader.data_load import DataLoader
    from oracle import io_layout, ref_loader
    rng = np.random.default_rng(12)
    nrow, ncol, ncell = 360, 720, 5000
    lin = rng.choice(nrow * ncol, ncell, replace=False)
    coords = np.zeros((ncell, 5))
    coords[:, 4], coords[:, 3] = lin % nrow + 1, lin // nrow + 1          # 1-based row / column (data_load.py:201-203)
    raster = rng.normal(500.0, 800.0, (280, ncol))                        # values below rep_val and "missing" cells
    raster[rng.random(raster.shape) < 0.05] = -9999
    dl = DataLoader.__new__(DataLoader)
    dl.s = SimpleNamespace(ncell=ncell, ngridrow=nrow, ngridcol=ncol)
    dl.coords = coords
    for rep in (None, 1000, 0):
        got = dl.load_routing_data(raster, rep_val=rep)
        want = io_layout.load_routing_vector(raster, coords, nrow, ncol, skip=68, rep_val=rep)
        assert got.shape == (ncell,) and bitwise_equal(got, want)
    if ref_loader.available():
        ref_loader.load()
        import xanthos.data_reader.data_load as rdl
        import xanthos.utils.math as rmath
        idx = rmath.sub2ind([nrow, ncol], coords[:, 4].astype(int) - 1, coords[:, 3].astype(int) - 1)
        assert np.array_equal(idx, io_layout.sub2ind([nrow, ncol], coords[:, 4].astype(int) - 1, coords[:, 3].astype(int) - 1))
        assert bitwise_equal(rdl.DataLoader.vectorize(raster, nrow, ncol, idx, skip=68),
                             io_layout.vectorize(raster, nrow, ncol, idx, 68))


def test_pacing_window_of_the_skew_kernel_is_deadlock_free(monkeypatch):
    """The month pacing window of `mrtm_skew_kernel` (a warp crosses boundary b only after every warp has crossed b - W)
    against the ring back-pressure and the chunked hand-over of the cut edges: the protocol is replayed on the plan
    tables (tests/skew_model.py::handover_completes).  With the window the library computes (`xan_mrtm_skew_window`:
    pipeline depth of the linked warps + ring capacity) every warp finishes - for 2 and 4 cells per lane, 3-hourly and
    hourly sub-steps, three networks including the 0.5 degree bench world - and a window shorter than the pipeline is
    shown to deadlock there (it did on the GPU: DESIGN.md section 4)."""
    import skew_model
    from xanthos_b200 import synthetic
    from xanthos_b200.routing import mrtm
    from oracle.calendar_utils import set_month_arrays
    worlds = [synthetic.make_world(40, 80, 1500, 8, seed=5, coast_pull=0.0), synthetic.make_world(90, 180, 6000, 30, seed=2),
              synthetic.make_world(seed=0)]
    nd = set_month_arrays(60, 2001, 2005)[:, 2]
    for w in worlds:
        s = w.settings()
        up = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
        for K in ('2', '4'):
            monkeypatch.setenv('XANTHOS_MRTM_SKEW_K', K)
            um = mrtm.upstream_genmatrix(up)
            t = skew_model.skew_tables(um)
            for dt, months, spin in ((10800, 48, 12), (3600, 14, 4)):
                nt = np.array([int(d * 24 * 3600 / dt) for d in list(nd[:spin]) + list(nd[:months])])
                win, ch, rl, dflt = skew_model.pacing_window(um, 1, nt.min())
                assert dflt >= 1 and win >= 3 and skew_model.pacing_window(um, 0, nt.min())[0] == 0
                assert skew_model.pacing_window(um, 10 ** 6, nt.min())[0] == 10 ** 6          # a larger request is kept
                assert skew_model.handover_completes(t, nt, win, ch, rl), (w.ncell, K, dt, win)
                assert skew_model.handover_completes(t, nt, 0, ch, rl)                         # pacing off
        if w.ncell > 60000:      # the outlets of the bench world run a dozen months behind the leaves
            nt = np.array([int(d * 24 * 3600 / 10800) for d in list(nd[:12]) + list(nd[:48])])
            assert t['info']['n_levels'] >= 10
            assert not skew_model.handover_completes(t, nt, 2, ch, rl)
