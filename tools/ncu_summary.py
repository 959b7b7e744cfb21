"""Summarise ncu CSV exports: (1) launch list -> time per kernel name; (2) raw pages -> key roofline metrics."""
import csv
import re
import sys
from collections import defaultdict


def launches(path):
    rows = [r for r in csv.reader(open(path, errors='ignore')) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    ui = hdr.index('Metric Unit')
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(',', ''))
        if r[ui] in ('ns', 'nsecond'):
            v /= 1e3
        elif r[ui] in ('ms', 'msecond'):
            v *= 1e3
        name = re.sub(r'\(.*', '', r[ki])
        name = re.sub(r'^void ', '', name)[:90]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f'total {tot / 1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised)')
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f'{us / 1e3:9.3f} ms {100 * us / tot:5.1f}% {n:5d}x  {k}')


def raw(path, keys):
    rows = list(csv.reader(open(path, errors='ignore')))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [h for h in hdr if any(k in h for k in keys)]
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[idx['Kernel Name']])[-60:]
        print(name, 'grid', r[idx.get('Grid Size', 0)], 'block', r[idx.get('Block Size', 0)])
        for c in cols:
            print(f'    {c:70s} {r[idx[c]]:>14s} {units[idx[c]]}')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3:])
