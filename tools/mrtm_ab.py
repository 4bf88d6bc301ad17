"""A/B timing of the MRTM warp kernel on the bench world (config 1: 360 spin-up + 360 months).

usage: python tools/mrtm_ab.py NAME=ENV1=v1,ENV2=v2 ...   e.g. tree=XANTHOS_MRTM_AUTO=tree skew=XANTHOS_MRTM_AUTO=skew   (each argument is one variant; the environment
variables are read by xan_mrtm_route at every call).  Prints the best and median of 3 runs per variant and
checks that every variant returns bit-identical ChStorage / Avg_ChFlow."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4

M = int(os.environ.get('MRTM_MONTHS', '360'))
w = synthetic.make_world(seed=0)
s = w.settings()
q = C.Field.from_host(synthetic.runoff_input(w, M, seed=3))
L, V, A = C.dev_vector(w.flow_dist), C.dev_vector(w.velocity), C.dev_vector(w.area)
nd = month_days_mod4(M, 1971)
up = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
ref = None
for arg in sys.argv[1:] or ['default=']:
    name, _, envs = arg.partition('=')
    kv = [e.split('=', 1) for e in envs.split(',') if e]
    for k, v in kv:
        os.environ[k] = v
    um = mrtm.upstream_genmatrix(up)   # the plan reads XANTHOS_MRTM_THREADS / _LANES when it is created
    if ref is None:
        print(um.info, flush=True)
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = mrtm.route_device(um, q, L, V, A, nd, 10800, M)
        e1.record()
        torch.cuda.synchronize()
        if it:
            ts.append(e0.elapsed_time(e1))
    sig = [o.t[:, :o.ncell].clone() for o in out[:2]]
    same = True if ref is None else all(torch.equal(a.view(torch.int64), b.view(torch.int64)) for a, b in zip(sig, ref))
    ref = ref or sig
    print('%-24s best %.2f ms  median %.2f ms  identical=%s' % (name, min(ts), float(np.median(ts)), same), flush=True)
    for k, v in kv:
        os.environ.pop(k, None)
