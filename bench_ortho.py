#!/usr/bin/env python
"""
bench_ortho.py — BASELINE.json configs[3]: orthoplane inference on a synthetic isotropic volume (default 1024^3)
feeding the RLE consensus, the flow of scripts/pdl_inference3d.py:110-240 with `-mode orthoplane`:

  for axis in xy, xz, yz:
      slices of the HBM-resident uint8 volume along the axis        inference/volume.take_slices (emp_take_slices)
      CNN forward per slice (stand-in net, see bench_stack.StandInNet; synthetic heads from the label volume added)
      stack post-processing + RLE tables for the whole axis         inference/stack.StackShard.finish
      forward + backward cross-slice matching                       StackShard.match (emp_rle_pair_overlaps + host Hungarian)
      tracker lifting 2D -> 3D                                      inference/tracker.InstanceTracker (host)
  instance consensus over the three trackers                        consensus.merge_objects_from_trackers
  dense fill of the consensus volume                                inference/fill (emp_fill_runs)

Prints one JSON line with the seconds of every stage.  One GPU; the three axes are independent stack passes and can be
given to three ranks (not done here).

    python bench_ortho.py [--size 1024] [--blobs 1500]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AXES = {'xy': 0, 'xz': 1, 'yz': 2}


def blob_volume(size, n_blobs, seed, dev, radii=(8.0, 22.0)):
    """(S,S,S) int32 label volume of axis-aligned ellipsoids painted on the device (later blobs overwrite earlier ones)."""
    import torch
    rng = np.random.default_rng(seed)
    vol = torch.zeros((size, size, size), dtype=torch.int32, device=dev)
    c = rng.uniform(0, size, (n_blobs, 3))
    rad = rng.uniform(radii[0], radii[1], (n_blobs, 3))
    for i in range(n_blobs):
        lo = np.maximum(np.floor(c[i] - rad[i]).astype(int), 0)
        hi = np.minimum(np.ceil(c[i] + rad[i]).astype(int) + 1, size)
        g = [torch.arange(lo[a], hi[a], device=dev, dtype=torch.float32) for a in range(3)]
        m = (((g[0] - c[i, 0]) / rad[i, 0]) ** 2)[:, None, None] + (((g[1] - c[i, 1]) / rad[i, 1]) ** 2)[None, :, None] + \
            (((g[2] - c[i, 2]) / rad[i, 2]) ** 2)[None, None, :] <= 1
        box = vol[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        box[m] = i + 1
    return vol


def heads_from_labels(lab, n_ids, gen, coarse=4, sigma=1.5):
    """Head tensors a trained model would emit for label slices `lab` (n, A, B) int32: full-res semantic probabilities
    (0.9 inside / 0.1 outside +- 0.05), quarter-res center heat-map (a Gaussian at every instance's centroid) and offsets
    (centroid - pixel, in full-res pixels) — the GT-style targets of data/utils/target_creation.py:13-78, on the device."""
    import torch
    import torch.nn.functional as F
    n, A, B = lab.shape
    dev = lab.device
    key = (torch.arange(n, device=dev)[:, None, None] * (n_ids + 1) + lab).reshape(-1).long()
    yy = torch.arange(A, device=dev, dtype=torch.float32)[None, :, None].expand(n, A, B).reshape(-1)
    xx = torch.arange(B, device=dev, dtype=torch.float32)[None, None, :].expand(n, A, B).reshape(-1)
    cnt = torch.zeros(n * (n_ids + 1), device=dev).index_add_(0, key, torch.ones_like(yy))
    cy = torch.zeros_like(cnt).index_add_(0, key, yy) / cnt.clamp(min=1)
    cx = torch.zeros_like(cnt).index_add_(0, key, xx) / cnt.clamp(min=1)
    sem = (0.1 + 0.8 * (lab > 0).float() + (torch.rand(lab.shape, device=dev, generator=gen) - 0.5) * 0.1)[:, None, None]
    a, b = A // coarse, B // coarse
    labc = lab[:, ::coarse, ::coarse][:, :a, :b]
    keyc = (torch.arange(n, device=dev)[:, None, None] * (n_ids + 1) + labc).long()
    thing = labc > 0
    Y = (torch.arange(a, device=dev, dtype=torch.float32) * coarse)[None, :, None]
    X = (torch.arange(b, device=dev, dtype=torch.float32) * coarse)[None, None, :]
    off = torch.stack([torch.where(thing, cy[keyc] - Y, torch.zeros((), device=dev)),
                       torch.where(thing, cx[keyc] - X, torch.zeros((), device=dev))], dim=1)[:, None]
    # heat-map: a unit impulse at every present instance's (coarse) centroid, blurred
    present = (cnt.view(n, n_ids + 1)[:, 1:] > 0).nonzero()
    hm = torch.zeros((n, 1, a, b), device=dev)
    if present.shape[0]:
        s_idx, k_idx = present[:, 0], present[:, 1] + 1
        flat = s_idx * (n_ids + 1) + k_idx
        py = (cy[flat] / coarse).round().clamp(0, a - 1).long()
        px = (cx[flat] / coarse).round().clamp(0, b - 1).long()
        hm[s_idx, 0, py, px] = 1.0
    r = int(3 * sigma)
    g1 = torch.exp(-(torch.arange(-r, r + 1, device=dev, dtype=torch.float32) ** 2) / (2 * sigma * sigma))
    hm = F.conv2d(F.conv2d(hm, g1.view(1, 1, -1, 1), padding=(r, 0)), g1.view(1, 1, 1, -1), padding=(0, r)).clamp(max=1.0)
    return sem.contiguous(), hm[:, None].contiguous(), off.contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--blobs', type=int, default=1500)
    ap.add_argument('--batch', type=int, default=64, help='slices taken from the volume per emp_take_slices call')
    ap.add_argument('--ks', type=int, default=3)
    args = ap.parse_args()
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    from empanada_b200.inference import patterns, stack, filters
    from empanada_b200.inference.volume import take_slices
    import bench_stack as bs

    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    S = args.size
    shape = (S, S, S)
    t0 = time.perf_counter()
    lab_vol = blob_volume(S, args.blobs, 7, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    img_vol = (100 + 60 * (lab_vol > 0).to(torch.float32) + 20 * torch.randn(shape, device=dev, generator=gen)).clamp(0, 255).to(torch.uint8)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    eng = bs.make_engine()
    net = bs.StandInNet(dev)
    L = eng.label_divisor
    trackers = patterns.create_axis_trackers(AXES, [1], L, shape)
    sem_b = torch.empty((S + args.ks // 2, 1, 1, S, S), dtype=torch.float32, device=dev)
    hm_b = torch.empty((S, 1, 1, S // 4, S // 4), dtype=torch.float32, device=dev)
    off_b = torch.empty((S, 1, 2, S // 4, S // 4), dtype=torch.float32, device=dev)
    # warm-up: library load, cuDNN autotune, allocator — a 4-slice stack along every axis
    for axis in AXES.values():
        w = stack.StackShard(eng, labels=[1], depth=4, median_kernel_size=args.ks)
        img = take_slices(img_vol, axis, 0, 4)
        sem, hm, off = heads_from_labels(take_slices(lab_vol, axis, 0, 4), args.blobs, gen)
        for s in range(4):
            net((img[s][None, None].to(torch.float32) - 130) / 38)
            w.add(s, sem[s], hm[s], off[s], size=(S, S))
        w.match(w.finish())
    torch.cuda.synchronize()
    per_axis = {}
    for name, axis in AXES.items():
        tm = {}
        shard = stack.StackShard(eng, labels=[1], depth=S, median_kernel_size=args.ks, upsampling=1, force_connected=True)
        t_extract = t_cnn = t_heads = 0.0
        for i0 in range(0, S, args.batch):
            n = min(args.batch, S - i0)
            torch.cuda.synchronize()
            t = time.perf_counter()
            img = take_slices(img_vol, axis, i0, n)
            torch.cuda.synchronize()
            t_extract += time.perf_counter() - t
            t = time.perf_counter()
            outs = []
            for s in range(n):
                image = (img[s][None, None].to(torch.float32) - 255 * 0.508979) / (255 * 0.148561)
                outs.append(net(image))
            torch.cuda.synchronize()
            t_cnn += time.perf_counter() - t
            t = time.perf_counter()
            lab = take_slices(lab_vol, axis, i0, n)
            sem, hm, off = heads_from_labels(lab, args.blobs, gen)
            for s in range(n):
                torch.add(sem[s], torch.sigmoid(outs[s]['sem_logits']), alpha=0.0, out=sem_b[i0 + s])
                torch.add(hm[s], outs[s]['ctr_hmp'], alpha=0.0, out=hm_b[i0 + s])
                torch.add(off[s], outs[s]['offsets'], alpha=0.0, out=off_b[i0 + s])
                shard.add(i0 + s, sem_b[i0 + s], hm_b[i0 + s], off_b[i0 + s], size=(S, S))
            torch.cuda.synchronize()
            t_heads += time.perf_counter() - t
        tm.update(take_slices_s=t_extract, cnn_s=t_cnn, synthetic_heads_s=t_heads)
        t = time.perf_counter()
        segs = shard.finish()
        torch.cuda.synchronize()
        tm['postproc_rle_s'] = time.perf_counter() - t
        n_inst, n_runs = segs.counts()
        t = time.perf_counter()
        matched = shard.match(segs)
        torch.cuda.synchronize()
        tm['match_s'] = time.perf_counter() - t
        t = time.perf_counter()
        for z in range(S):
            patterns.update_trackers(matched[z], z, trackers[name])
        patterns.finish_tracking(trackers[name])
        tm['tracker_s'] = time.perf_counter() - t
        tm.update(instances_2d=n_inst, runs_2d=n_runs, objects_3d=len(trackers[name][0].instances))
        per_axis[name] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in tm.items()}
        print(f'[ortho] {name}: {per_axis[name]}', file=sys.stderr, flush=True)
        del shard, segs, matched
    t = time.perf_counter()
    cons = patterns.create_instance_consensus(patterns.get_axis_trackers_by_class(trackers, 1), 2, 0.75, False)
    t_cons = time.perf_counter() - t
    t = time.perf_counter()
    out_vol = torch.zeros(shape, dtype=torch.int32, device=dev)
    patterns.fill_volume(out_vol, cons.instances)
    torch.cuda.synchronize()
    t_fill = time.perf_counter() - t
    # how well the consensus reproduces the ground truth it was rendered from (a sanity figure, not a parity claim)
    agree = float(((out_vol > 0) == (lab_vol > 0)).float().mean().item())
    gpu_s = sum(a['take_slices_s'] + a['postproc_rle_s'] for a in per_axis.values())
    rec = {'metric': 'orthoplane_postproc', 'config': {'workload': f'orthoplane_{S}^3', 'blobs': args.blobs, 'median_kernel_size': args.ks,
                                                        'data': 'synthetic blob volume; stand-in CNN + GT-style heads'},
           'setup_s': round(t_setup, 3), 'per_axis': per_axis, 'consensus_s': round(t_cons, 4), 'fill_s': round(t_fill, 4),
           'consensus_objects': len(cons.instances), 'foreground_agreement_with_ground_truth': round(agree, 5),
           'voxels_per_s_take_plus_postproc_rle': 3 * S ** 3 / gpu_s,
           'note': 'take_slices + postproc_rle are this library\'s kernels; match is one overlap launch + the host Hungarian chain; '
                   'tracker and consensus are host code as in the reference (consensus overlaps on the GPU)'}
    os.write(json_fd, (json.dumps(rec) + '\n').encode())


if __name__ == '__main__':
    main()
