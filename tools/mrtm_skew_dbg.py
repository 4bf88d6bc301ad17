"""Per-warp cycle accounting of the skew kernel (XANTHOS_MRTM_DEBUG): total / prologue wait / events."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4
M = int(os.environ.get('MRTM_MONTHS', '120'))
w = synthetic.make_world(seed=0)
s = w.settings()
q = C.Field.from_host(synthetic.runoff_input(w, M, seed=3))
L, V, A = C.dev_vector(w.flow_dist), C.dev_vector(w.velocity), C.dev_vector(w.area)
nd = month_days_mod4(M, 1971)
um = mrtm.upstream_genmatrix(mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s))
mrtm.route_device(um, q, L, V, A, nd, 10800, M, method=C.MRTM_SKEW)
torch.cuda.synchronize()
f = 'gpurun_out/skew_dbg.txt'
os.environ['XANTHOS_MRTM_DEBUG'] = f
mrtm.route_device(um, q, L, V, A, nd, 10800, M, method=C.MRTM_SKEW)
torch.cuda.synchronize()
d = np.loadtxt(f)
T = sum(int(x) * 8 for x in nd) * 2
tot, wait, evt, sp, dw, nslow = d[:, 1], d[:, 2], d[:, 3], d[:, 4], d[:, 5], d[:, 6]
print('iterations', T, 'warps', len(d))
print('total cycles/iter   pct 0/50/90/100:', np.percentile(tot / T, [0, 50, 90, 100]).round(1))
print('wait  cycles/iter   pct 0/50/90/100:', np.percentile(wait / T, [0, 50, 90, 100]).round(1))
print('event cycles/iter   pct 0/50/90/100:', np.percentile(evt / T, [0, 50, 90, 100]).round(1))
loop = (tot - wait - evt) / T
print('loop  cycles/iter   pct 0/50/90/100:', np.percentile(loop, [0, 50, 90, 100]).round(1))
free = d[:, 2] == 0
print('free warps', free.sum(), 'loop median', np.median(loop[free]).round(1), 'linked loop median', np.median(loop[~free]).round(1))
print('slow iterations fraction pct 0/50/90/100:', np.percentile(nslow / T, [0, 50, 90, 100]).round(3), 'mean', (nslow / T).mean().round(3))
cnt = np.bincount(sp.astype(int) // 4, minlength=148)
print('warps per SM min/max', cnt.min(), cnt.max())
for k in (0, 10, 20, 28):
    m = (dw >= k) & (dw < k + 10)
    if m.any():
        print('Dw %d..%d: n=%d total/iter %.1f evt/iter %.1f' % (k, k + 9, m.sum(), np.median(tot[m] / T), np.median(evt[m] / T)))
