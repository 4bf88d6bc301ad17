"""Time the MRTM warp kernel for several (block threads, chunk, lanes) choices on the bench world."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4

w = synthetic.make_world(seed=0)
s = w.settings()
M = int(os.environ.get('MRTM_MONTHS', '360'))
q = C.Field.from_host(synthetic.runoff_input(w, M, seed=3))
L, V, A = C.dev_vector(w.flow_dist), C.dev_vector(w.velocity), C.dev_vector(w.area)
nd = month_days_mod4(M, 1971)
up = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
ref = None
cfgs = [(256, 64, 32), (128, 64, 32), (64, 64, 32), (256, 32, 32), (256, 124, 32), (256, 248, 32), (256, 64, 24), (256, 64, 16)]
if len(sys.argv) > 1:
    cfgs = [tuple(int(x) for x in a.split(',')) for a in sys.argv[1:]]
for (T, CH, lanes) in cfgs:
    os.environ['XANTHOS_MRTM_LANES'] = str(lanes)
    um = mrtm.upstream_genmatrix(up, T, CH)
    info = um.info
    try:
        mrtm.route_device(um, q, L, V, A, nd, 10800, 12)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        chs, avg, inst = mrtm.route_device(um, q, L, V, A, nd, 10800, M)
        e1.record()
        torch.cuda.synchronize()
        sig = float(avg.t[:, :w.ncell].sum())
        if ref is None:
            ref = sig
        print("T=%d chunk=%d lanes=%d warps=%d edges=%d levels=%d  %.2f ms  same=%s" % (T, CH, lanes, info['n_warps'], info['n_cut_edges'], info['n_levels'], e0.elapsed_time(e1), sig == ref), flush=True)
    except Exception as ex:
        print("T=%d chunk=%d lanes=%d failed: %s" % (T, CH, lanes, ex), flush=True)
