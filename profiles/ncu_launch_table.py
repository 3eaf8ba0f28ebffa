#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file RAW)
-> a per-launch csv and a per-kernel summary (count, total / mean time, DRAM read / write bytes per launch).

    python profiles/ncu_launch_table.py RAW OUT.csv [--skip N] [--title "..."]

Every launch of the run is listed (ours and torch's), so the share of our kernels in a step can be read off."""
import argparse
import collections
import csv
import re

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3,
        'msecond': 1e3, 'second': 1e6}


def short(name):
    name = re.sub(r'^(void )?', '', name)
    name = re.sub(r'\(.*$', '', name)
    name = name.replace('emp::', '')
    return name[:90]


def load(raw):
    rows = [r for r in csv.reader(l for l in open(raw, errors='replace') if l.startswith('"'))]
    col = {h: i for i, h in enumerate(rows[0])}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[col['ID']], {'kernel': short(r[col['Kernel Name']]), 'grid': r[col['Grid Size']], 'block': r[col['Block Size']]})
        d[r[col['Metric Name']]] = float(r[col['Metric Value']].replace(',', '')) * UNIT.get(r[col['Metric Unit']], 1.0)
    return list(launches.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('raw')
    ap.add_argument('out')
    ap.add_argument('--skip', type=int, default=0, help='launches to leave out at the start (warm-up)')
    ap.add_argument('--title', default='')
    a = ap.parse_args()
    L = load(a.raw)[a.skip:]
    agg = collections.OrderedDict()
    for d in L:
        g = agg.setdefault(d['kernel'], {'n': 0, 't': 0.0, 'r': 0.0, 'w': 0.0})
        g['n'] += 1
        g['t'] += d.get('gpu__time_duration.sum', 0.0)
        g['r'] += d.get('dram__bytes_read.sum', 0.0)
        g['w'] += d.get('dram__bytes_write.sum', 0.0)
    total = sum(g['t'] for g in agg.values()) or 1.0
    with open(a.out, 'w') as f:
        f.write(f'# {a.title}\n# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none '
                f'(cold-cache, serialised launches: read SHARES, not absolutes)\n')
        f.write('# ---- per kernel: launches, total us, share, mean us, DRAM read MB / launch, DRAM write MB / launch\n')
        for k, g in sorted(agg.items(), key=lambda kv: -kv[1]['t']):
            f.write(f'# {k},{g["n"]},{g["t"]:.1f},{g["t"] / total:.3f},{g["t"] / g["n"]:.2f},{g["r"] / g["n"] / 1e6:.3f},{g["w"] / g["n"] / 1e6:.3f}\n')
        f.write('id,kernel,grid,block,time_us,dram_read_MB,dram_write_MB\n')
        for i, d in enumerate(L):
            f.write(f'{i},{d["kernel"]},"{d["grid"]}","{d["block"]}",{d.get("gpu__time_duration.sum", 0):.2f},'
                    f'{d.get("dram__bytes_read.sum", 0) / 1e6:.3f},{d.get("dram__bytes_write.sum", 0) / 1e6:.3f}\n')
    for k, g in sorted(agg.items(), key=lambda kv: -kv[1]['t'])[:25]:
        print(f'{k[:70]:70s} n={g["n"]:4d} total={g["t"]:9.1f}us share={g["t"] / total:.3f} mean={g["t"] / g["n"]:8.2f}us '
              f'rd={g["r"] / g["n"] / 1e6:8.2f}MB wr={g["w"] / g["n"] / 1e6:8.2f}MB')


if __name__ == '__main__':
    main()
