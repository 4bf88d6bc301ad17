"""ncu report(s) -> profiles JSON: per kernel DRAM bytes per launch, duration, FP64-pipe / issue utilisation, warps active.
usage: python tools/ncu_summary.py OUT.json REPORT.ncu-rep [REPORT2.ncu-rep ...]   (first launch of every kernel name wins)"""
import csv, io, json, subprocess, sys

WANT = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active': 'fp64_pipe_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct_of_ncu_peak',
    'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
    'launch__registers_per_thread': 'registers',
    'smsp__inst_executed.sum': 'warp_instructions',
}
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'usecond': 1e-3,
         'msecond': 1.0, 'nsecond': 1e-6, 'second': 1e3}


def short(name):
    n = name.replace('void ', '').replace('xan::', '')
    head = n.split('(')[0]
    return head.replace('(bool)', '').replace('(int)', '').replace(' ', '')


out = {'source': ', '.join(sys.argv[2:]) + ' (ncu --set full --clock-control none; first launch of every kernel)', 'kernels': {}}
for rep in sys.argv[2:]:
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        k = short(r[ix['Kernel Name']])
        if k in out['kernels']:
            continue
        d = {}
        for m, key in WANT.items():
            if m in ix and r[ix[m]] != '':
                v = float(r[ix[m]].replace(',', ''))
                d[key] = v * SCALE.get(units[ix[m]], 1)
        if 'dram_read' in d:
            d['dram_bytes_per_launch'] = d.pop('dram_read') + d.pop('dram_write', 0.0)
        d['ms_under_ncu'] = d.pop('duration', None)
        out['kernels'][k] = d
json.dump(out, open(sys.argv[1], 'w'), indent=1)
print(json.dumps({k: (round(v.get('ms_under_ncu') or 0, 4), round(v.get('dram_bytes_per_launch', 0) / 1e6, 1)) for k, v in out['kernels'].items()}))
