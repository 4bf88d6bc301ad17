"""MRTM throughput with 1, 2, 4 ensemble members per call (two members per warp pass) on the bench world."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4

w = synthetic.make_world(seed=0)
s = w.settings()
M = 360
qs = [C.Field.from_host(synthetic.runoff_input(w, M, seed=3 + k)) for k in range(4)]
nd = month_days_mod4(M, 1971)
um = mrtm.upstream_genmatrix(mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s))
for k in (1, 2, 4):
    mrtm.route_device_batch(um, qs[:k], w.flow_dist, w.velocity, w.area, nd, 10800, 12)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mrtm.route_device_batch(um, qs[:k], w.flow_dist, w.velocity, w.area, nd, 10800, M)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('%d member(s): %.2f ms, %.2f ms per member, %.3e cell-months/s (routing only, with spin-up)'
          % (k, ms, ms / k, k * w.ncell * M / (ms * 1e-3)), flush=True)
