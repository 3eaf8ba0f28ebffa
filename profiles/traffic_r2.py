#!/usr/bin/env python
"""profiles/traffic.json from the round's ncu launch lists: DRAM bytes per pixel of the config-2 kernels (bench.py reads
them for roofline.traffic) and DRAM bytes per voxel of the stack path's kernels (one launch over a 64-slice block)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ncu_launch_table import load  # noqa: E402

ROOT = os.path.dirname(HERE)


def biggest(launches, prefix):
    """the launch of kernel `prefix` that moved the most bytes (the full-batch one)"""
    c = [d for d in launches if d['kernel'].startswith(prefix)]
    return max(c, key=lambda d: d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) if c else None


def main():
    tiles = load(os.path.join(ROOT, 'gpurun_out', 'r2_launches_raw.csv'))
    allk = load(os.path.join(ROOT, 'gpurun_out', 'r2_all_raw.csv'))
    px = 16 * 4096 * 4096
    out = {'source': 'profiles/r2_launches.csv / r2_all_kernels.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch)',
           'dram_bytes_per_px': {}, 'ncu_time_us': {}, 'stack_dram_bytes_per_voxel': {}, 'stack_ncu_time_us_per_64_slices': {}}
    for key, prefix in (('nms_peaks', 'nms_peaks_kernel'), ('assign', 'assign_kernel<1, 0, 0, 1>'), ('apply_lut', 'apply_lut_staged_kernel')):
        d = biggest(tiles, prefix)
        out['dram_bytes_per_px'][key] = (d['dram__bytes_read.sum'] + d['dram__bytes_write.sum']) / px
        out['ncu_time_us'][key] = d['gpu__time_duration.sum']
    vox = 64 * 2048 * 2048
    for prefix in ('median_chain_kernel', 'merge_lean_kernel', 'rle_block_emit_kernel', 'assign_kernel<2, 0, 3, 1>',
                   'rle_block_union_kernel', 'rle_block_flags_kernel', 'rle_block_slots_kernel', 'rle_block_assign_kernel',
                   'rle_block_offsets_kernel', 'rle_block_lists_kernel', 'rle_block_pack_kernel', 'rle_block_keys_kernel'):
        d = biggest(allk, prefix)
        if d:
            out['stack_dram_bytes_per_voxel'][prefix] = (d['dram__bytes_read.sum'] + d['dram__bytes_write.sum']) / vox
            out['stack_ncu_time_us_per_64_slices'][prefix] = d['gpu__time_duration.sum']
    with open(os.path.join(HERE, 'traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
