"""Per-warp cycle accounting (XANTHOS_MRTM_DEBUG dump) grouped by SM sub-partition."""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1])
nsub = float(sys.argv[2]) if len(sys.argv) > 2 else 175328.
tot, wait, stage, loop, redo, sp = a[:, 1] / nsub, a[:, 2] / nsub, a[:, 3] / nsub, a[:, 4] / nsub, a[:, 5] / nsub, a[:, 6].astype(int)
print(len(a), 'warps; total max %.0f mean %.0f cycles/sub-step' % (tot.max(), tot.mean()))
print('loop: mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f' % (loop.mean(), np.median(loop), np.percentile(loop, 90), np.percentile(loop, 99), loop.max()))
r = redo > 0.1
print('redo warps %d: loop mean %.0f max %.0f | others: loop mean %.0f max %.0f' % (r.sum(), loop[r].mean(), loop[r].max(), loop[~r].mean(), loop[~r].max()))
cnt = np.bincount(sp)
print('warps per sub-partition (histogram):', np.bincount(cnt[cnt > 0]).tolist(), ' redo warps per sub-partition:', np.bincount(np.bincount(sp[r], minlength=len(cnt))).tolist())
for i in np.argsort(-loop)[:10]:
    print('  warp %4d loop %.0f wait %.0f stage %.0f tot %.0f redo %.2f sp %d co-residents %s' % (a[i, 0], loop[i], wait[i], stage[i], tot[i], redo[i], sp[i], np.round(loop[sp == sp[i]]).astype(int).tolist()))
