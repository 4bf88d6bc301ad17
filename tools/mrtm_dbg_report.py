"""Per-warp cycle accounting of the MRTM warp kernel (XANTHOS_MRTM_DEBUG=<file>) by warp category."""
import sys
import numpy as np
d = np.load(sys.argv[2]) if len(sys.argv) > 2 else None
a = np.loadtxt(sys.argv[1])
nsub = float(sys.argv[3]) if len(sys.argv) > 3 else 175328.
tot, wait, stage, loop = a[:, 1] / nsub, a[:, 2] / nsub, a[:, 3] / nsub, a[:, 4] / nsub
print('warps %d  per sub-step cycles: total mean %.0f max %.0f | loop mean %.0f p50 %.0f p90 %.0f max %.0f | wait mean %.0f | stage mean %.0f max %.0f'
      % (len(a), tot.mean(), tot.max(), loop.mean(), np.median(loop), np.percentile(loop, 90), loop.max(), wait.mean(), stage.mean(), stage.max()))
if d is not None:
    nw = len(a)
    ngh = np.bincount(d['ec'], minlength=nw)
    nout = np.bincount(d['ep'], minlength=nw)
    for g in range(0, 5):
        for o in (0, 1):
            m = (ngh == g) & ((nout > 0) == bool(o))
            if m.any():
                print('  ghosts %d out %d: n=%4d loop %.0f wait %.0f stage %.0f other %.0f' % (g, o, m.sum(), loop[m].mean(), wait[m].mean(), stage[m].mean(), (tot - loop - wait - stage)[m].mean()))
