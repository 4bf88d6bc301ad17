"""Where does the full-size ABCD run differ from the oracle?  (run on a B200; prints the worst cells)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xanthos_b200 import synthetic
from xanthos_b200.runoff import abcd
from oracle import abcd as oabcd
w = synthetic.make_world(seed=0)
m = 360
ab = synthetic.abcd_inputs(w, m, seed=1)
tmin = np.nan_to_num(ab['tmin'])
pet, aet, q, sav = abcd.abcd_execute(w.n_basins, w.basin_ids, ab['pet'], ab['precip'], tmin, ab['pars'], m, 360, -1)
want = oabcd.abcd_execute(w.n_basins, w.basin_ids, ab['pet'], ab['precip'], tmin, ab['pars'], m, 360)
for got, ref, name in ((aet, want[1], 'aet'), (q, want[2], 'q'), (sav, want[3], 'sav')):
    ok = ~np.isnan(ref)
    err = np.zeros(ref.shape)
    err[ok] = np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-6)
    print(name, 'max err', err.max(), 'cells > 1e-9:', int((err.max(axis=1) > 1e-9).sum()), 'entries > 1e-9:', int((err > 1e-9).sum()))
    order = np.argsort(-err.max(axis=1))[:6]
    for c in order:
        k = int(np.argmax(err[c]))
        first = int(np.argmax(err[c] > 1e-12))
        b = int(w.basin_ids[c])
        print('  cell', c, 'basin', b, 'pars', ab['pars'][b - 1], 'month', k, 'got', got[c, k], 'ref', ref[c, k], 'err', err[c, k],
              'first month > 1e-12:', first, 'pet/precip there', ab['pet'][c, k], ab['precip'][c, k])
    # per basin: how many cells are off
    bad = err.max(axis=1) > 1e-9
    print('  basins with bad cells:', np.unique(w.basin_ids[bad])[:20], 'of cells', bad.sum())
