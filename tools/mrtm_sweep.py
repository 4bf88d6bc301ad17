"""Time the MRTM tree kernel for several (threads, cells/thread, block fill) choices on the bench world."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4

w = synthetic.make_world(seed=0)
s = w.settings()
M = 360
q = C.Field.from_host(synthetic.runoff_input(w, M, seed=3))
L, V, A = C.dev_vector(w.flow_dist), C.dev_vector(w.velocity), C.dev_vector(w.area)
nd = month_days_mod4(M, 1971)
up = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
ref = None
for (T, K, fill) in [(256, 2, 0), (512, 1, 0), (256, 1, 0), (128, 2, 0), (128, 1, 0), (256, 2, 300), (512, 2, 0), (128, 4, 0), (64, 2, 0)]:
    if fill:
        os.environ['XANTHOS_MRTM_BLOCK_CELLS'] = str(fill)
    else:
        os.environ.pop('XANTHOS_MRTM_BLOCK_CELLS', None)
    um = mrtm.upstream_genmatrix(up, T, K)
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
        print("T=%d K=%d fill=%d blocks=%d edges=%d levels=%d  %.2f ms  same=%s" % (T, K, fill, info['n_blocks'], info['n_cut_edges'], info['n_levels'], e0.elapsed_time(e1), sig == ref), flush=True)
    except Exception as ex:
        print("T=%d K=%d fill=%d failed: %s" % (T, K, fill, ex), flush=True)
