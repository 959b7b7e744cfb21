"""Summarise an ncu launch list (CSV from `ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,dram__bytes_*`):
per kernel name -> launches, device time and share of the step, time-weighted tensor-pipe activity, DRAM bytes.
   python tools/ncu_launch_summary.py gpurun_out/r2_launches.csv > profiles/r2_ncu_launches_....summary.txt"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path, errors='ignore')) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ui, idi = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
    per = defaultdict(dict)  # launch id -> metric -> value
    names = {}
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        u = r[ui]
        if r[mi].startswith('gpu__time_duration'):
            v = v / 1e3 if u in ('ns', 'nsecond') else v * 1e3 if u in ('ms', 'msecond') else v  # -> us
        if r[mi].startswith('dram__bytes'):
            v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        per[r[idi]][r[mi]] = v
        names[r[idi]] = re.sub(r'^void ', '', re.sub(r'\(.*', '', r[ki]))[:80]
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])  # n, us, tensor%*us, dram bytes
    for i, m in per.items():
        us = m.get('gpu__time_duration.sum', 0.0)
        a = agg[names[i]]
        a[0] += 1
        a[1] += us
        a[2] += us * m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0)
        a[3] += m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0)
    tot = sum(a[1] for a in agg.values())
    own = sum(a[1] for k, a in agg.items() if 'sb::' in k)
    print(f'{sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms of device time (cold-cache, serialised under ncu: '
          f'compare SHARES); kernels of this library: {100 * own / tot:.1f} % of it')
    print('   time ms  share   n   tensor-pipe active %   DRAM MB   kernel')
    for k, (n, us, tw, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f'{us / 1e3:9.3f} {100 * us / tot:5.1f}% {n:4d}   {tw / us if us else 0:8.1f}            {by / 1e6:9.1f}   {k}')


if __name__ == '__main__':
    main(sys.argv[1])
