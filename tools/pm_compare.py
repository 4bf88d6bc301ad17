"""Penman-Monteith: throughput kernel vs exact-order kernel on the full bench workload (time + max deviation)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.pet import penman_monteith as pm

w = synthetic.make_world(seed=0)
d = synthetic.pm_inputs(w, 1971, 2000, seed=1)
names = ('tair_load', 'TMIN_load', 'rhs_load', 'wind_load', 'rsds_load', 'rlds_load')
for k in names:
    d[k] = C.Field.from_host(np.nan_to_num(d[k]))
d['lct_load'] = pm.stage_land_cover(d['lct_load'], d['tair_load'].ld)
d['elev'] = C.dev_vector(d['elev'])
ns = SimpleNamespace(**d)
out = {}
for mode in ('1', '0'):
    os.environ['XANTHOS_PM_EXACT'] = mode
    for _ in range(2):
        f = pm.run_pmpet_device(ns, w.ncell, d['nlcs'], 1971, 2000, d['water_idx'], d['snow_idx'], d['lc_years'])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f = pm.run_pmpet_device(ns, w.ncell, d['nlcs'], 1971, 2000, d['water_idx'], d['snow_idx'], d['lc_years'])
    e1.record()
    torch.cuda.synchronize()
    out[mode] = f.t[:, :w.ncell].clone()
    print('exact' if mode == '1' else 'fast ', '%.3f ms' % (e0.elapsed_time(e1) / 5))
a, b = out['1'], out['0']
rel = ((a - b).abs() / a.abs().clamp_min(1e-6)).max().item()
print('max relative deviation fast vs exact over %d cell-months: %.3e' % (a.numel(), rel), ' nan mismatch', int((torch.isnan(a) != torch.isnan(b)).sum()))
absd = (a - b).abs()
i = int(torch.argmax(absd / a.abs().clamp_min(1e-6)))
m, c = divmod(i, a.shape[1])
print('worst: month %d cell %d exact %.17g fast %.17g abs diff %.3e' % (m, c, a[m, c].item(), b[m, c].item(), absd[m, c].item()))
print('max abs deviation %.3e mm/month; median PET %.3f' % (absd.max().item(), a.median().item()))
for floor in (1e-6, 1e-4, 1e-2, 1.0):
    print('  max rel with floor %g: %.3e' % (floor, (absd / a.abs().clamp_min(floor)).max().item()))
# both kernels against the numpy oracle on the first two years, all cells
from oracle import pet as opet
dd = synthetic.pm_inputs(w, 1971, 1972, seed=1)
for k in names:
    dd[k] = np.nan_to_num(dd[k])
want = opet.pm_pet(dd, w.ncell, dd['nlcs'], 1971, 1972, dd['water_idx'], dd['snow_idx'], dd['lc_years'])
ns2 = SimpleNamespace(**dd)
for mode in ('1', '0'):
    os.environ['XANTHOS_PM_EXACT'] = mode
    got = pm.run_pmpet(ns2, w.ncell, dd['nlcs'], 1971, 1972, dd['water_idx'], dd['snow_idx'], dd['lc_years'])
    ad = np.abs(got - want)
    print('exact' if mode == '1' else 'fast ', 'vs oracle (67,420 x 24): max abs %.3e' % ad.max(),
          ' max rel floor 1e-6: %.3e' % (ad / np.maximum(np.abs(want), 1e-6)).max(),
          ' floor 1e-3: %.3e' % (ad / np.maximum(np.abs(want), 1e-3)).max())
