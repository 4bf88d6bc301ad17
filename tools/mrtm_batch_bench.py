"""MRTM throughput with several ensemble members per launch of the skew kernel (their blocks share the SMs) on the
bench world: K cells per lane x members per launch, every variant checked bit-identical to one launch per member."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4

w = synthetic.make_world(seed=0)
s = w.settings()
M = int(os.environ.get('MONTHS', 360))
DT = float(os.environ.get('DT', 10800))
NMEM = 6
qs = [C.Field.from_host(synthetic.runoff_input(w, M, seed=3 + k)) for k in range(NMEM)]
nd = month_days_mod4(M, 1971)
upid = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
ref = None
for K in [int(v) for v in os.environ.get('KS', '2,4').split(',')]:
    os.environ['XANTHOS_MRTM_SKEW_K'] = os.environ['XANTHOS_MRTM_SKEW_KM'] = str(K)
    um = mrtm.upstream_genmatrix(upid)
    for nm, win in [(int(a), int(b)) for a in os.environ.get('NMS', '1,2,3,4').split(',') for b in os.environ.get('WINDOWS', '0').split(',')]:
        os.environ['XANTHOS_MRTM_SKEW_MEMBERS'] = str(nm)
        os.environ['XANTHOS_MRTM_SKEW_WINDOW'] = str(win)
        mrtm.route_device_batch(um, qs, w.flow_dist, w.velocity, w.area, nd, DT, 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = mrtm.route_device_batch(um, qs, w.flow_dist, w.velocity, w.area, nd, DT, M)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if ref is None:
            ref = [(o[0].t.clone(), o[1].t.clone(), o[2].clone()) for o in outs]
        same = all(torch.equal(o[0].t.view(torch.int64), r[0].view(torch.int64)) and
                   torch.equal(o[1].t.view(torch.int64), r[1].view(torch.int64)) and
                   torch.equal(o[2].view(torch.int64), r[2].view(torch.int64)) for o, r in zip(outs, ref))
        print('K=%d window=%d members per launch=%d: %d members %.2f ms, %.2f ms per member, %.3e cell-months/s, bitwise same: %s'
              % (K, win, nm, NMEM, ms, ms / NMEM, NMEM * w.ncell * M / (ms * 1e-3), same), flush=True)
        del outs
