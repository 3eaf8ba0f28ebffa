#!/usr/bin/env python
"""Every kernel of libempanada_b200 once (twice) at BASELINE sizes — the command the per-kernel ncu launch list of the
round is taken from:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
        --log-file gpurun_out/all_raw.csv python profiles/all_kernels.py
    python profiles/ncu_launch_table.py gpurun_out/all_raw.csv profiles/r2_all_kernels.csv --title "..."

Sections: config 2 (16 x 4096^2 fused batch), the reference-shaped standalone functions and emp_rle on a 4096^2 tile,
emp_median_harden / Render path / emp_rle on a 2048^2 slice, a 64-slice stack block (emp_median_chain, emp_stack_blocks),
the matcher's overlap kernel and the dense fill over that block, the consensus overlap kernel."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from empanada_b200 import _cabi as C                                          # noqa: E402
from empanada_b200 import consensus                                           # noqa: E402
from empanada_b200.inference import engines, postprocess as pp, rle, stack    # noqa: E402
from empanada_b200.synth import synth_tile                                    # noqa: E402
import bench_stack as bs                                                      # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    reps = 2
    # ---- config 2: the fused batch
    H = W = 4096
    B = 16
    tiles = [synth_tile(H, W, 500, seed=b) for b in range(B)]
    sem = torch.stack([torch.from_numpy(t['sem'][0, 0]) for t in tiles]).to(dev)
    hm = torch.stack([torch.from_numpy(t['ctr_hmp'][0, 0]) for t in tiles]).to(dev)
    off = torch.stack([torch.from_numpy(t['offsets'][0]) for t in tiles]).to(dev)
    L = C.lib()
    things, nt = C.i64_array([1])
    k_cap = pp.DEFAULT_K_CAP
    per_tile = L.emp_workspace_bytes(H, W, k_cap, nt)
    ws = torch.empty(per_tile * B, dtype=torch.uint8, device=dev)
    pan = torch.empty((B, H, W), dtype=torch.int64, device=dev)
    for _ in range(reps + 1):
        C.check(L.emp_panoptic_batched(B, sem.data_ptr(), 0, hm.data_ptr(), off.data_ptr(), H, W, things, nt, 1000, 64, 0, 0.1, 7,
                                       pan.data_ptr(), None, 0, k_cap, ws.data_ptr(), per_tile, C.stream_ptr(dev)))
    torch.cuda.synchronize()
    # ---- one tile through the reference-shaped functions, then RLE
    s1, h1, o1 = sem[:1, None], hm[:1, None], off[:1]
    for _ in range(reps):
        ctr = pp.find_instance_center(h1, 0.1, 7)
        ids = pp.group_pixels(ctr, o1)
        ins, _ = pp.get_instance_segmentation(s1[0], h1, o1, [1], 0.1, 7)
        merged = pp.merge_semantic_and_instance(s1[0], ins, 1000, [1], 64, 0)
        rle.rle_tables(merged[0].contiguous(), [1], 1000, [1], True)
    del sem, hm, off, pan, ws
    # ---- config 3: one 2048^2 slice the per-slice way, then a 64-slice block the batched way
    slices = bs.make_slices(dev, 2048, 12, 400)
    eng = bs.make_engine()
    planes = [s['sem_prob'] for s in slices[:3]]
    for _ in range(reps):
        engines.median_harden(planes, 0.3, want_median=True, want_sem='u8')
        p3 = eng._fused_postprocess(planes[1], slices[1]['ctr_hmp'], slices[1]['offsets'], 1)
        rle.rle_tables(p3[0], [1], 20000, [1], True)
    shard = None
    for _ in range(reps):
        shard = stack.StackShard(eng, labels=[1], depth=64, median_kernel_size=3, block=64)
        for z in shard.slices():
            s = slices[z % len(slices)]
            shard.add(z, s['sem_prob'], s['ctr_hmp'], s['offsets'], size=(2048, 2048))
        out = shard.finish()
    matched = shard.match(out)                          # rle_pair_overlaps over the block's run tables
    vol = shard.fill(torch.int64)                       # fill_runs over the block
    del vol
    # ---- consensus: overlaps between the objects of two "trackers" (here: two shifted copies of a slice's instances)
    seg = out[5][1]
    starts = [a['starts'] for a in seg.values()] + [a['starts'] + 3 for a in seg.values()]
    runs = [a['runs'] for a in seg.values()] * 2
    tracker = [0] * len(seg) + [1] * len(seg)
    consensus.object_overlaps(tracker, starts, runs, dev)
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
