#!/usr/bin/env python
"""
bench_stack.py — BASELINE.json configs[2]: the 3D stack post-processing path on a synthetic
512 x 2048 x 2048 anisotropic volume, z-sharded over N GPUs of one node, RLE out.

What is timed (per rank, max over ranks): everything between "the CNN heads of my z-block are in HBM"
and "every slice's RLE tables are on the host with globally unique labels" — the recursive median
chain + harden (engines.py:47-90,114-121), coarse center search + group_pixels(step=4) on the
512 x 512 maps (:257-272), the upsample-fused merge (:274-292), pan_seg -> RLE (rle.py:26-86),
the D2H of the run tables, the carry-plane exchange between neighbouring ranks and the all-gather of
instance counts (inference/stack.py).  The CNN forward is the unchanged PyTorch/cuDNN model and is
not part of this path (its heads are synthetic here, cycled from a few distinct slices; slice z is
the same on every rank, so a sharded run can be compared with a single block).

    python bench_stack.py [--depth 512] [--hw 2048] [--ks 3]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_stack.py --gpus N

Prints one JSON line (rank 0): voxels/s over all ranks, ms per slice, scaling "strong" (the volume is
fixed, ranks split it).  bench.py runs the same function at every N and puts the record on its line.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_VOXEL = 20.75     # sem prob 4 + (ctr 4 + off 8) / 16 + pan 8 (post-proc) + pan 8 (RLE read)


def make_slices(dev, hw, distinct=12, blobs=400, seed=100):
    """`distinct` head-tensor slices of a synthetic blob volume (slices 40.. of it, so blobs are live)."""
    import torch
    from empanada_b200.synth import synth_stack_slices
    slices = []
    for i, s in enumerate(synth_stack_slices(40 + distinct, hw, hw, blobs, seed=seed, coarse=4)):
        if i >= 40:
            slices.append({k: torch.from_numpy(s[k]).to(dev) for k in ('sem_prob', 'ctr_hmp', 'offsets')})
    return slices


def make_engine():
    """pdl_inference3d.py defaults (:28-37): nms kernel 3, threshold 0.1, confidence 0.3, label divisor 20000"""
    import torch
    from empanada_b200.inference import engines
    return engines.PanopticDeepLabRenderEngine(torch.nn.Identity(), thing_list=[1], label_divisor=20000, stuff_area=64, void_label=0,
                                               nms_threshold=0.1, nms_kernel=3, confidence_thr=0.3, coarse_boundaries=True)


def slice_digests(out):
    """{z: digest} over everything a slice's RLE dict holds (labels before the rank offset)."""
    dig = {}
    for z in out:
        rows = out.inst_rows(z)
        h = hashlib.blake2b(digest_size=12)
        if rows is None:                                    # a slice redone synchronously: hash the dict
            for c, d in out[z].items():
                for lab, a in d.items():
                    h.update(np.asarray([c, lab, *a['box']], np.int64).tobytes())
                    h.update(np.ascontiguousarray(a['starts']).tobytes())
                    h.update(np.ascontiguousarray(a['runs']).tobytes())
        else:
            t, _ = out._where[z]
            h.update(np.ascontiguousarray(rows[:, [0, 1, 2, 3, 4, 5, 6, 8]]).tobytes())
            for first, cnt in zip(rows[:, 7].tolist(), rows[:, 6].tolist()):
                h.update(t.starts[first:first + cnt].tobytes())
                h.update(t.lens[first:first + cnt].tobytes())
        dig[int(z)] = h.hexdigest()
    return dig


def run_stack(dev, rank, world, slices, depth, hw, ks=3, block=0, chain_chunk=4096, repeats=5, warmup=3, group=None,
              profile=False, match=False):
    """Times StackShard.finish() on this rank's z-block; returns (record, last RleStack, shard, matched)."""
    import torch
    import torch.distributed as dist
    from empanada_b200.inference import stack
    eng = make_engine()
    if not block:                                            # short blocks: smaller sub-blocks keep the copy / parse pipeline busy
        block = 128                                          # StackShard clips it to the block's length

    # The heads of this rank's block (+ halo) as the inference loop leaves them: one batch buffer per head, slice z a
    # view into it (run_stack_with_cnn shows the CNN writing them there) — StackShard then passes base + stride on.
    z0, _ = stack.partition_slices(depth, world, rank)
    _, zh = stack.halo_range(depth, world, rank, ks)
    idx = torch.tensor([z % len(slices) for z in range(z0, zh)], device=dev)
    sem_b = torch.stack([s['sem_prob'] for s in slices]).index_select(0, idx)
    hm_b = torch.stack([s['ctr_hmp'] for s in slices]).index_select(0, idx)
    off_b = torch.stack([s['offsets'] for s in slices]).index_select(0, idx)

    def run_once():
        shard = stack.StackShard(eng, labels=[1], depth=depth, rank=rank, world_size=world, median_kernel_size=ks,
                                 upsampling=1, force_connected=True, block=block, chain_chunk=chain_chunk, group=group)
        for i, z in enumerate(shard.slices()):
            shard.add(z, sem_b[i], hm_b[i], off_b[i], size=(hw, hw))
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(group=group)
            torch.cuda.synchronize(dev)
        t = time.perf_counter()
        out = shard.finish()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t
        if world > 1:                                       # the job is done when its slowest rank is
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=group)
            dt = float(tt.item())
        return dt, out, shard

    for _ in range(warmup):
        run_once()
    times, out, shard = [], None, None
    for _ in range(repeats):
        dt, out, shard = run_once()
        times.append(dt)
    n_inst, n_runs = out.counts()
    timing = dict(getattr(shard, 'timing_', {}))
    t = time.perf_counter()
    out.materialise()                                       # every slice's nested dict (what a dict-walking consumer pays on top)
    timing['dicts_s'] = time.perf_counter() - t
    matched = None
    if match:
        t = time.perf_counter()
        matched = shard.match(out)
        torch.cuda.synchronize(dev)
        timing['match_s'] = time.perf_counter() - t
    stages = None
    if profile:
        from empanada_b200 import _cabi as C
        C.profile_enable(True)
        run_once()
        stages = {k: [round(v[0], 4), v[1]] for k, v in C.profile_read().items() if v[1]}
        C.profile_enable(False)
    per_rank = None
    if world > 1:
        c = torch.tensor([n_inst, n_runs], device=dev, dtype=torch.int64)
        dist.all_reduce(c, group=group)
        n_inst, n_runs = int(c[0]), int(c[1])
        gathered = [None] * world                           # the last repeat's host timeline of every rank
        dist.all_gather_object(gathered, {k: round(v, 5) for k, v in timing.items()}, group=group)
        per_rank = gathered
    sec = sorted(times)[len(times) // 2]                     # median: the caching allocator takes a few blocks to settle
    rec = {'metric': 'stack_postproc_throughput', 'value': depth * hw * hw / sec, 'unit': 'voxels/s', 'n_gpus': world,
           'seconds': sec, 'seconds_is': 'median of the repeats, max over ranks each', 'seconds_best': min(times), 'seconds_all': [round(t, 5) for t in times], 'repeats': repeats, 'ms_per_slice_per_rank': 1e3 * sec / max(len(out), 1),
           'scaling': 'strong', 'rank0_host_seconds': {k: round(v, 4) for k, v in timing.items()},
           'per_rank_host_seconds': per_rank, 'stage_ms_and_launches': stages,
           'config': {'workload': f'stack_{depth}x{hw}x{hw}_coarse4_ks{ks}', 'slices_per_rank': len(out), 'block': block,
                      'instances': n_inst, 'rle_runs': n_runs, 'data': 'synthetic head tensors, CNN not included'}}
    return rec, out, shard, matched


class StandInNet:
    """A small conv net with PanopticDeepLab's output contract (full-res semantic logits, quarter-res center heat-map and
    offsets: quantization/panoptic_deeplab.py:238-250 with coarse boundaries): the unchanged-CNN side of the stack loop.
    Random-init heads are constants, so the synthetic head tensors are ADDED to its outputs — the forward
    still has to run, the post-processing still sees EM-shaped data."""

    def __init__(self, dev, width=64):
        import torch
        import torch.nn as nn
        torch.manual_seed(0)
        self.body = nn.Sequential(
            nn.Conv2d(1, width // 2, 3, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(width // 2, width, 3, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(width, width, 3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(width, width, 3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(width, width, 3, padding=1), nn.ReLU(inplace=True)).to(dev).eval().to(memory_format=torch.channels_last)
        self.heads = nn.Conv2d(width, 4, 1).to(dev).eval()          # sem logit, center, offset y, offset x at 1/4 resolution
        torch.backends.cudnn.benchmark = True

    def __call__(self, image):
        import torch
        import torch.nn.functional as F
        with torch.no_grad():
            x = self.heads(self.body(image.contiguous(memory_format=torch.channels_last)))
            sem_logits = F.interpolate(x[:, 0:1], scale_factor=4, mode='bilinear', align_corners=True)
            return {'sem_logits': sem_logits, 'ctr_hmp': x[:, 1:2], 'offsets': x[:, 2:4]}


def run_stack_with_cnn(dev, rank, world, slices, depth, hw, ks=3, repeats=3, group=None):
    """The stack loop of scripts/pdl_inference3d.py:163-187 on this rank's z-block with the CNN in it: per slice a
    forward of the stand-in net on a uint8 image (normalised as VolumeDataset does), its heads written straight into
    the block's batch buffers (StackShard takes views: no copy), then the block's post-processing.  Timed twice:
    CNN alone, and CNN + post-processing + RLE tables on the host."""
    import torch
    import torch.distributed as dist
    from empanada_b200.inference import stack
    eng = make_engine()
    net = StandInNet(dev)
    z0, z1 = stack.partition_slices(depth, world, rank)
    _, zh = stack.halo_range(depth, world, rank, ks)
    n = zh - z0
    vol = torch.randint(0, 256, (n, 1, 1, hw, hw), dtype=torch.uint8, device=dev)
    sem_b = torch.empty((n, 1, 1, hw, hw), dtype=torch.float32, device=dev)
    hm_b = torch.empty((n, 1, 1, hw // 4, hw // 4), dtype=torch.float32, device=dev)
    off_b = torch.empty((n, 1, 2, hw // 4, hw // 4), dtype=torch.float32, device=dev)

    side = torch.cuda.Stream(dev)

    def run_once(postproc, streamed=False):
        shard = stack.StackShard(eng, labels=[1], depth=depth, rank=rank, world_size=world, median_kernel_size=ks,
                                 upsampling=1, force_connected=True, group=group, keep_tables=False,
                                 stream=side if streamed else None)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(group=group)
            torch.cuda.synchronize(dev)
        t = time.perf_counter()
        for i, z in enumerate(shard.slices()):
            image = (vol[i].to(torch.float32) - 255 * 0.508979) / (255 * 0.148561)       # A.Normalize of the MitoNet configs
            o = net(image)
            s = slices[z % len(slices)]
            torch.add(s['sem_prob'], torch.sigmoid(o['sem_logits']), alpha=0.0, out=sem_b[i])
            torch.add(s['ctr_hmp'], o['ctr_hmp'], alpha=0.0, out=hm_b[i])
            torch.add(s['offsets'], o['offsets'], alpha=0.0, out=off_b[i])
            shard.add(z, sem_b[i], hm_b[i], off_b[i], size=(hw, hw))
            if streamed and (i + 1) % 16 == 0:
                shard.advance()                                 # complete sub-blocks leave on the side stream, beside the CNN
        out = shard.finish() if postproc else None
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=group)
            dt = float(tt.item())
        return dt, out

    run_once(True)                                              # warm-up (cudnn autotune, workspaces, allocator)
    run_once(True)
    t_cnn = min(run_once(False)[0] for _ in range(repeats))
    both = [run_once(True) for _ in range(repeats)]
    t_all = min(b[0] for b in both)
    n_inst, n_runs = both[-1][1].counts()
    # the same with StackShard.advance(): sub-blocks of 128 slices are chained and cut into tables on a side stream while the
    # CNN works on the following slices (ranks whose block is a single sub-block have nothing to stream)
    run_once(True, True)
    st = [run_once(True, True) for _ in range(repeats)]
    t_streamed = min(b[0] for b in st)
    same = slice_digests(st[-1][1]) == slice_digests(both[-1][1])
    return {'metric': 'stack_inference_throughput', 'value': depth * hw * hw / t_all, 'unit': 'voxels/s', 'n_gpus': world,
            'seconds': t_all, 'seconds_cnn_only': t_cnn, 'postproc_share_of_wall': (t_all - t_cnn) / t_all, 'scaling': 'strong',
            'seconds_streamed': t_streamed, 'streamed_postproc_share_of_wall': (t_streamed - t_cnn) / t_streamed,
            'streamed_equals_block': bool(same),
            'config': {'workload': f'stack_{depth}x{hw}x{hw}_coarse4_ks{ks}_with_cnn', 'slices_per_rank_incl_halo': n,
                       'cnn': 'stand-in conv net (5 conv layers, 64 channels, quarter-res heads, bilinear x4 semantic logits), fp32 — two orders of magnitude lighter than the ResNet-50 PanopticDeepLab it stands for',
                       'instances_rank0': n_inst, 'rle_runs_rank0': n_runs,
                       'data': 'uint8 noise volume through the net; synthetic head tensors added to its (constant-like) outputs'}}


def add_parity(rec, out, dev, rank, world, slices, depth, hw, ks, block, chain_chunk, group=None):
    """N > 1: every rank's slices must equal the same stack run as ONE block (rank 0 runs and times it)."""
    import torch.distributed as dist
    mine = slice_digests(out)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    if rank == 0:
        one, whole, _, _ = run_stack(dev, 0, 1, slices, depth, hw, ks, block, chain_chunk, 5, warmup=3)
        want = slice_digests(whole)
        got = {}
        for g in gathered:
            got.update(g)
        rec['parity'] = 'bit-equal to single-block on rank 0' if got == want else 'MISMATCH vs single-block on rank 0'
        rec['single_gpu_seconds_same_run'] = one['seconds']
        rec['speedup_vs_n1'] = one['seconds'] / rec['seconds']
    dist.barrier(group=group)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--depth', type=int, default=512)
    ap.add_argument('--hw', type=int, default=2048)
    ap.add_argument('--ks', type=int, default=3)
    ap.add_argument('--distinct', type=int, default=12, help='distinct synthetic slices (cycled)')
    ap.add_argument('--blobs', type=int, default=400)
    ap.add_argument('--repeat', type=int, default=5)
    ap.add_argument('--block', type=int, default=0, help='slices per emp_stack_block call (0: 128)')
    ap.add_argument('--chain-chunk', type=int, default=4096, help='slices per emp_median_chain launch')
    ap.add_argument('--profile', action='store_true', help='one extra run with per-stage CUDA events (ms per stage over the block)')
    ap.add_argument('--cnn', action='store_true', help='also time the stack loop with a stand-in CNN in it')
    ap.add_argument('--match', action='store_true', help='also time the cross-slice matcher (forward + backward) on the block')
    ap.add_argument('--match-cpu-slices', type=int, default=12, help='slices of the CPU matcher baseline (oracle port of the reference)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    sys.stdout.flush()
    json_fd = os.dup(1)                 # stdout carries the ONE JSON line; NCCL's banner etc. go to stderr
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        try:                                                # every rank on its own slice of the host's CPUs
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            os.sched_setaffinity(0, cpus[local_rank * per:(local_rank + 1) * per] or cpus)
        except OSError:
            pass
    D, H = args.depth, args.hw
    t0 = time.time()
    slices = make_slices(dev, H, args.distinct, args.blobs)
    if rank == 0:
        print(f'[rank 0] {len(slices)} distinct slices ready in {time.time() - t0:.1f} s', file=sys.stderr, flush=True)
    rec, out, shard, matched = run_stack(dev, rank, world, slices, D, H, args.ks, args.block, args.chain_chunk, args.repeat,
                                         profile=args.profile, match=args.match)
    if world > 1:
        add_parity(rec, out, dev, rank, world, slices, D, H, args.ks, args.block, args.chain_chunk)
    match_cpu = None
    if args.match and rank == 0:
        # the reference's matcher on the host (oracle port, numpy): forward chain over the first slices
        from oracle import matcher as om
        zs = sorted(out)[:args.match_cpu_slices]
        t = time.perf_counter()
        m = om.RLEMatcher(1, 20000, 0.25, 0.25, True)
        for z in zs:
            seg = out[z][1]
            if m.target_rle is None:
                m.initialize_target(seg)
            else:
                m(seg)
        match_cpu = {'ms_per_slice_forward_only': 1e3 * (time.perf_counter() - t) / max(len(zs) - 1, 1), 'slices': len(zs),
                     'kind': 'port', 'cores': 1}
    if args.cnn:
        del out, shard, matched
        with_cnn = run_stack_with_cnn(dev, rank, world, slices, D, H, args.ks)
        rec['with_cnn'] = with_cnn
    if rank == 0:
        rec['matcher_cpu_baseline'] = match_cpu
        os.write(json_fd, (json.dumps(rec) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
