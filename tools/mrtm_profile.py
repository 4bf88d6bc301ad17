"""Run the MRTM tree kernel once on the bench world for a few months (target of an ncu capture)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4

M = int(os.environ.get('MRTM_MONTHS', '24'))
w = synthetic.make_world(seed=0)
s = w.settings()
q = C.Field.from_host(synthetic.runoff_input(w, M, seed=3))
L, V, A = C.dev_vector(w.flow_dist), C.dev_vector(w.velocity), C.dev_vector(w.area)
nd = month_days_mod4(M, 1971)
up = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
um = mrtm.upstream_genmatrix(up)
print(um.info)
for _ in range(2):
    mrtm.route_device(um, q, L, V, A, nd, 10800, 0)
torch.cuda.synchronize()
