"""Latency of ONE warp: a tiny world whose cells fit a single warp."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4
for (nr, nc, n) in [(12, 24, 28), (12, 24, 60), (16, 32, 200)]:
    w = synthetic.make_world(nr, nc, n, 1, seed=1, coast_pull=0.0, edge_cases=False)
    s = w.settings()
    M = 360
    q = C.Field.from_host(synthetic.runoff_input(w, M, seed=3))
    L, V, A = C.dev_vector(w.flow_dist), C.dev_vector(w.velocity), C.dev_vector(w.area)
    nd = month_days_mod4(M, 1971)
    up = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
    um = mrtm.upstream_genmatrix(up)
    mrtm.route_device(um, q, L, V, A, nd, 10800, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mrtm.route_device(um, q, L, V, A, nd, 10800, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nsub = int(nd.sum()) * 8
    print(um.info, 'max nup', int(up[:, 8].max()), '%.2f ms  %.0f ns/sub-step = %.0f cycles' % (ms, ms * 1e6 / nsub, ms * 1e6 / nsub * 1.965))
