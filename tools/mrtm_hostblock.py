"""Does xan_mrtm_route return before the kernel has run?  Host time per call of a back-to-back series
(no synchronisation in between) against the device time per call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4

M = 360
w = synthetic.make_world(seed=0)
s = w.settings()
q = C.Field.from_host(synthetic.runoff_input(w, M, seed=3))
L, V, A = C.dev_vector(w.flow_dist), C.dev_vector(w.velocity), C.dev_vector(w.area)
nd = month_days_mod4(M, 1971)
up = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
for arg in sys.argv[1:] or ['default=']:
    name, _, envs = arg.partition('=')
    kv = [e.split('=', 1) for e in envs.split(',') if e]
    for k, v in kv:
        os.environ[k] = v
    um = mrtm.upstream_genmatrix(up)
    mrtm.route_device(um, q, L, V, A, nd, 10800, M)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = []
    e0.record()
    for _ in range(5):
        t0 = time.perf_counter()
        mrtm.route_device(um, q, L, V, A, nd, 10800, M)
        host.append((time.perf_counter() - t0) * 1e3)
    e1.record()
    torch.cuda.synchronize()
    print('%-16s device %.2f ms per call; host time per call %s' % (name, e0.elapsed_time(e1) / 5, np.round(host, 2).tolist()), flush=True)
    for k, v in kv:
        os.environ.pop(k, None)
