"""Per-warp cycle accounting (XANTHOS_MRTM_DEBUG, member 0) of a two-member launch of the skew kernel, with the block ->
warp-set map of the second member rotated or not: who are the warps that never wait (they set the pace)?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from bench import month_days_mod4
M = int(os.environ.get('MRTM_MONTHS', '120'))
w = synthetic.make_world(seed=0)
s = w.settings()
qs = [C.Field.from_host(synthetic.runoff_input(w, M, seed=3 + k)) for k in range(2)]
L, V, A = C.dev_vector(w.flow_dist), C.dev_vector(w.velocity), C.dev_vector(w.area)
nd = month_days_mod4(M, 1971)
um = mrtm.upstream_genmatrix(mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s))
T = sum(int(x) * 8 for x in nd) * 2
for rot in os.environ.get('ROTS', '0,74').split(','):
    os.environ['XANTHOS_MRTM_SKEW_ROTATE'] = rot
    os.environ.pop('XANTHOS_MRTM_DEBUG', None)
    mrtm.route_device_batch(um, qs, L, V, A, nd, 10800, M)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mrtm.route_device_batch(um, qs, L, V, A, nd, 10800, M)
    e1.record()
    torch.cuda.synchronize()
    f = 'gpurun_out/pair_dbg_%s.txt' % rot
    os.environ['XANTHOS_MRTM_DEBUG'] = f
    mrtm.route_device_batch(um, qs, L, V, A, nd, 10800, M)
    torch.cuda.synchronize()
    d = np.loadtxt(f)
    tot, wait, evt, sp, dw, nslow = d[:, 1], d[:, 2], d[:, 3], d[:, 4], d[:, 5], d[:, 6]
    busy = (tot - wait) / T
    print('rotate %s: %.2f ms for 2 members x %d months; %d warps; total cycles/iter max %.1f' % (rot, e0.elapsed_time(e1), M, len(d), (tot / T).max()))
    print('  busy cycles/iter pct 0/50/90/99/100:', np.percentile(busy, [0, 50, 90, 99, 100]).round(1))
    print('  slow fraction    pct 0/50/90/99/100:', np.percentile(nslow / T, [0, 50, 90, 99, 100]).round(3))
    order = np.argsort(-busy)[:12]
    for i in order:
        print('   warp %4d busy %.1f wait %.1f evt %.1f slow %.3f Dw %d sm %d sp %d' % (d[i, 0], busy[i], wait[i] / T, evt[i] / T, nslow[i] / T, dw[i], sp[i] // 4, sp[i] % 4))
    print('  corr(busy, slow fraction) = %.3f' % np.corrcoef(busy, nslow / T)[0, 1])
