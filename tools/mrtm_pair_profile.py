"""Target of an ncu capture: the skew kernel on the bench world, one member per launch (two launches), then two members in
one launch (MRTM_MONTHS months, no spin-up)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4

M = int(os.environ.get('MRTM_MONTHS', '48'))
w = synthetic.make_world(seed=0)
s = w.settings()
NMS = [int(v) for v in os.environ.get('NMS', '1,2').split(',')]
qs = [C.Field.from_host(synthetic.runoff_input(w, M, seed=3 + k)) for k in range(max(NMS))]
nd = month_days_mod4(M, 1971)
um = mrtm.upstream_genmatrix(mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s))
for nm in NMS:
    os.environ['XANTHOS_MRTM_SKEW_MEMBERS'] = str(nm)
    mrtm.route_device_batch(um, qs[:nm], w.flow_dist, w.velocity, w.area, nd, 10800, 0)
    torch.cuda.synchronize()
