"""Post-processing scans and the DE generation kernels once each at the bench size - target of an ncu capture."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.drought import drought_stats as dr
from xanthos_b200.diagnostics import time_series as ts
from xanthos_b200.accessible import accessible as acc

w = synthetic.make_world(seed=0)
n, m = w.ncell, 360
rng = np.random.default_rng(5)
q = C.Field.from_host(np.abs(rng.normal(50.0, 40.0, (n, m))))
for _ in range(2):
    thr = dr.getthresh_device(q.t, n, 12)
    S, I, D = dr.droughtstats_device(q.t, n, thr)
    agg = ts.group_sum_device(w.basin_ids, q.t)
    ann = acc.basin_annual_runoff_device(q, w.area, w.basin_ids)
# DE kernels on the calibration shape: 235 problems x 75 members x 5 parameters
nb, S_, D_ = 235, 75, 5
dev = dict(dtype=torch.float64, device='cuda')
pop = torch.empty((nb, S_, D_), **dev)
lib = C.lib()
C.check(lib.xan_de_init(C.ptr(pop), nb, S_, D_, 7, C.stream_ptr()))
E = ((pop - 0.3) ** 2).sum(dim=2).contiguous()
act = torch.arange(nb, dtype=torch.int32, device='cuda')
lo = torch.full((D_,), 1e-4, **dev)
span = torch.full((D_,), 0.9998, **dev)
tx, tp = torch.empty((nb, S_, D_), **dev), torch.empty((nb, S_, 5), **dev)
conv = torch.zeros(nb, dtype=torch.int32, device='cuda')
for g in (1, 2):
    C.check(lib.xan_de_trial(C.ptr(pop), C.ptr(E), C.ptr(act), nb, S_, D_, 5, C.ptr(lo), C.ptr(span), 7, g, 0.5, 1.0, 0.7,
                             C.ptr(tx), C.ptr(tp), C.stream_ptr()))
    Et = ((tx - 0.3) ** 2).sum(dim=2).contiguous()
    C.check(lib.xan_de_select(C.ptr(pop), C.ptr(E), C.ptr(act), nb, S_, D_, C.ptr(tx), C.ptr(Et), 0.01, 0.0, g, C.ptr(conv),
                              C.stream_ptr()))
torch.cuda.synchronize()
print('ok')
