"""Throughput of the streamflow calibration objective (StreamflowEvaluator) on the bench world: all 235 basins x P candidate
parameter sets per generation = P global ABCD runs + P / 2 two-member routing launches (360 + 360 months)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from xanthos_b200.calibrate import calibrate_abcd as cal
from bench import month_days_mod4

M, P = 360, int(os.environ.get('POP', 8))
w = synthetic.make_world(seed=0)
s = w.settings()
ab = synthetic.abcd_inputs(w, M, seed=1)
dsid = mrtm.downstream(w.coords, w.flow_dir, s)
um = mrtm.upstream_genmatrix(mrtm.upstream(w.coords, dsid, s))
ev = cal.StreamflowEvaluator(w.basin_ids, w.area, C.Field.from_host(ab['precip']), C.Field.from_host(ab['pet']),
                             C.Field.from_host(np.nan_to_num(ab['tmin'])), M, M, um, dsid, w.flow_dist, w.velocity,
                             month_days_mod4(M, 1971), 10800, M)
rng = np.random.default_rng(4)
bn = np.arange(1, w.n_basins + 1)
lo, hi = np.array([b[0] for b in cal.BOUNDS_SNOW]), np.array([b[1] for b in cal.BOUNDS_SNOW])
cand = lo + (hi - lo) * rng.random((w.n_basins, P, 5))
_, series = ev.evaluate(bn, np.broadcast_to(ab['pars'][:, None, :], (w.n_basins, 1, 5)).copy(), np.ones((w.n_basins, M)),
                        want_series=True)
obs = series[:, 0, :] * (1 + rng.normal(0, 0.05, (w.n_basins, M)))
ev.evaluate(bn, cand[:, :2], obs)
torch.cuda.synchronize()
t0 = time.perf_counter()
ed = ev.evaluate(bn, cand, obs)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print('streamflow objective: %d basins x %d candidates in %.1f ms = %.3e param-sets/s (%.1f ms per population slot); '
      'finite distances: %d of %d' % (w.n_basins, P, dt * 1e3, w.n_basins * P / dt, dt * 1e3 / P, np.isfinite(ed).sum(), ed.size))
