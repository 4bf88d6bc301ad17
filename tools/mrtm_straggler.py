"""Isolated cost of a warp that repeats the balance at every other sub-step (a cell with dt * V / L > 1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4
for nfast in (0, 1, 2, 8):
    w = synthetic.make_world(16, 32, 200, 1, seed=1, coast_pull=0.0, edge_cases=False)
    s = w.settings()
    M = 360
    ds = mrtm.downstream(w.coords, w.flow_dir, s)
    up = mrtm.upstream(w.coords, ds, s)
    L, V = w.flow_dist.copy(), w.velocity.copy()
    inner = np.nonzero((up[:, 8] >= 1) & (ds > 0))[0]            # cells with both an upstream and a downstream cell
    for c in inner[:nfast]:
        L[c], V[c] = 1000.0, 1.0
    q = C.Field.from_host(synthetic.runoff_input(w, M, seed=3))
    nd = month_days_mod4(M, 1971)
    um = mrtm.upstream_genmatrix(up)
    dbg = '/tmp/straggler_%d.txt' % nfast
    os.environ['XANTHOS_MRTM_DEBUG'] = dbg
    mrtm.route_device(um, q, L, V, w.area, nd, 10800, 0)
    torch.cuda.synchronize()
    a = np.loadtxt(dbg).reshape(-1, 7)
    nsub = float(nd.sum() * 8)
    print('fast cells %d: warps %d; per warp loop cycles/sub-step %s redo %s' % (
        nfast, len(a), (a[:, 4] / nsub).astype(int).tolist(), np.round(a[:, 5] / nsub, 2).tolist()), flush=True)
