"""One differential-evolution generation of the calibration config (235 basins x 64 candidates) - for ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.calibrate import calibrate_abcd as cal
w = synthetic.make_world(seed=0)
M = 360
ab = synthetic.abcd_inputs(w, M, seed=2, with_pet=True)
ev = cal.BasinEvaluator(w.basin_ids, w.area, ab['precip'], ab['pet'], np.nan_to_num(ab['tmin']), M, M, 'km3_per_mth')
rng = np.random.default_rng(4)
P = int(os.environ.get('CALIB_POP', '64'))
lo = np.array([b[0] for b in cal.BOUNDS_SNOW]); hi = np.array([b[1] for b in cal.BOUNDS_SNOW])
pars = lo + (hi - lo) * rng.random((w.n_basins, P, 5))
obs = rng.uniform(0.5, 2.0, (w.n_basins, M))
bn = np.arange(1, w.n_basins + 1)
for _ in range(2):
    ev.evaluate(bn, pars, obs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ev.evaluate(bn, pars, obs)
e1.record()
torch.cuda.synchronize()
print('generation: %.2f ms -> %.0f param-sets/s' % (e0.elapsed_time(e1), w.n_basins * P / e0.elapsed_time(e1) * 1e3))
