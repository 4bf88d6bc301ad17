"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file RAW.csv) -> per-kernel summary CSV.
usage: python tools/launch_list_summary.py RAW.csv OUT.csv "header comment" """
import csv, re, sys
from collections import OrderedDict

raw, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
rows = [r for r in csv.reader(l for l in open(raw) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}
agg = OrderedDict()
for r in rows[1:]:
    try:
        ms = float(r[iv].replace(',', '')) * scale.get(r[iu], 1e-6)
    except ValueError:
        continue
    k = re.sub(r'\(.*$', '', r[ik].replace('void ', '').replace('xan::', '')).replace('(int)', '').replace('(bool)', '')
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(out, 'w') as f:
    if note:
        f.write('# %s\n' % note)
    f.write('kernel,launches,total_ms,share_pct,ms_per_launch\n')
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.4f,%.2f,%.4f\n' % (k, n, ms, 100 * ms / tot, ms / n))
print(open(out).read()[:1500])
