#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X)
-> profiles/r1_launches.csv (one row per emp:: launch) and profiles/traffic.json (DRAM bytes per pixel and time
share per kernel, read by bench.py for roofline.traffic).

    python profiles/launch_summary.py gpurun_out/launches_raw.csv [--tiles 16] [--hw 4096]
"""
import argparse
import collections
import csv
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3,
        'msecond': 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('raw')
    ap.add_argument('--tiles', type=int, default=16)
    ap.add_argument('--hw', type=int, default=4096)
    ap.add_argument('--command', default='python bench.py --tiles 16 --steps 2 --warmup 3 --no-cpu-baseline')
    args = ap.parse_args()
    rows = [r for r in csv.reader(l for l in open(args.raw) if l.startswith('"'))]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        name = r[col['Kernel Name']]
        if 'emp::' not in name and not re.search(r'(nms_peaks|emit_centers|bin_centers|assign|build_lut|apply_lut)', name):
            continue
        d = launches.setdefault(r[col['ID']], {'kernel': re.sub(r'^(void )?(emp::)?', '', name).split('(')[0],
                                               'grid': r[col['Grid Size']], 'block': r[col['Block Size']]})
        val = float(r[col['Metric Value']].replace(',', '')) * UNIT.get(r[col['Metric Unit']], 1.0)
        d[r[col['Metric Name']]] = val
    out = os.path.join(HERE, 'r1_launches.csv')
    with open(out, 'w') as f:
        f.write(f'# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none  {args.command}\n')
        f.write('# (cold-cache, serialised: compare SHARES with bench.py stage_ms_per_step, not absolutes); every emp:: launch of the run\n')
        f.write('id,kernel,grid,block,time_us,dram_read_MB,dram_write_MB\n')
        for i, d in enumerate(launches.values()):
            f.write(f'{i},{d["kernel"]},"{d["grid"]}","{d["block"]}",{d["gpu__time_duration.sum"]:.2f},'
                    f'{d["dram__bytes_read.sum"] / 1e6:.2f},{d["dram__bytes_write.sum"] / 1e6:.2f}\n')
    # per kernel: the last FULL-BATCH launch (the run ends with per-tile launches of the host-buffer path: those move
    # 1/tiles of the bytes and are left out, as are kernel variants that only the per-tile path uses)
    def nbytes(d):
        return d['dram__bytes_read.sum'] + d['dram__bytes_write.sum']
    biggest = {}
    for d in launches.values():
        biggest[d['kernel']] = max(biggest.get(d['kernel'], 0.0), nbytes(d))
    last = collections.OrderedDict()
    for d in launches.values():
        if nbytes(d) >= 0.9 * biggest[d['kernel']] and not d['kernel'].startswith('assign_kernel<2'):
            last[d['kernel']] = d
    px = args.tiles * args.hw * args.hw
    total = sum(d['gpu__time_duration.sum'] for d in last.values())
    key = {'nms_peaks': 'nms_peaks', 'assign': 'assign_kernel', 'apply_lut': 'apply_lut'}
    traffic = {'source': f'profiles/r1_launches.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch over {args.tiles} x {args.hw}^2 px, {args.command})',
               'dram_bytes_per_px': {k: next((d['dram__bytes_read.sum'] + d['dram__bytes_write.sum']) / px for n, d in last.items() if n.startswith(v))
                                     for k, v in key.items()},
               'ncu_time_us': {n: d['gpu__time_duration.sum'] for n, d in last.items()},
               'ncu_time_share': {n: d['gpu__time_duration.sum'] / total for n, d in last.items()}}
    with open(os.path.join(HERE, 'traffic.json'), 'w') as f:
        json.dump(traffic, f, indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == '__main__':
    main()
