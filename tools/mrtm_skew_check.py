"""Skew kernel (method = MRTM_SKEW) against the oracle, bit for bit: small worlds, then the bench world.
usage: python tools/mrtm_skew_check.py [full_months]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from xanthos_b200 import synthetic, _cuda as C
from xanthos_b200.routing import mrtm
from oracle import mrtm as om
from oracle.calendar_utils import set_month_arrays
from util import bitwise_equal


def check(w, months, spin, dt, tag):
    s = w.settings()
    up = mrtm.upstream(w.coords, mrtm.downstream(w.coords, w.flow_dir, s), s)
    um = mrtm.upstream_genmatrix(up)
    q = synthetic.runoff_input(w, months, seed=3)
    nd = set_month_arrays(24, 1971, 1972)[:months, 2]
    t0 = time.time()
    got = mrtm.route(um, q, w.flow_dist, w.velocity, w.area, nd, dt, spin, method=C.MRTM_SKEW)
    t1 = time.time()
    oup = om.upstream_fast(w.coords, om.downstream(w.coords, w.flow_dir, w.nrow, w.ncol), w.nrow, w.ncol)
    want = om.route(q, w.flow_dist, w.velocity, w.area, nd, dt, om.csr_rows(oup), spin)
    ok = [bool(bitwise_equal(a, b)) for a, b in zip(got, want)]
    print(tag, ok, 'gpu %.2fs' % (t1 - t0), flush=True)
    if not all(ok):
        for a, b, name in zip(got, want, ('chs', 'avg', 'inst')):
            bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
            print(name, 'mismatches', len(bad), bad[:5].tolist(), flush=True)
    return all(ok)


good = True
for (nr, nc, ncell, nb, seed) in [(24, 48, 320, 5, 43), (36, 72, 1500, 12, 0)]:
    w = synthetic.make_world(nr, nc, ncell, nb, seed=seed)
    for dt, months, spin in ((10800.0, 5, 2), (21600.0, 3, 3), (3600.0, 2, 1)):
        good &= check(w, months, spin, dt, 'world %d dt %g' % (ncell, dt))
fm = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if fm:
    w = synthetic.make_world(seed=0)
    good &= check(w, fm, min(fm, 6), 10800.0, 'bench world %d months' % fm)
print('ALL OK' if good else 'FAILED')
sys.exit(0 if good else 1)
