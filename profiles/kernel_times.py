#!/usr/bin/env python
"""Device time (CUDA events, median of N) of the path's other entry points at BASELINE sizes:
median+harden, the Render-engine coarse path, pan_seg -> RLE, and the standalone reference-shaped
functions.  Prints one JSON object; run on a B200 box:  python profiles/kernel_times.py"""
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from empanada_b200 import _cabi as C                                  # noqa: E402
from empanada_b200.inference import engines, postprocess as pp, rle   # noqa: E402
from empanada_b200.synth import synth_stack_slices, synth_tile        # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def stages(fn, n=10):
    """per-kernel device time through the library's own event pairs"""
    fn()
    torch.cuda.synchronize()
    C.profile_enable(True)
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    prof = C.profile_read()
    C.profile_enable(False)
    return {k: round(v[0] / n, 4) for k, v in prof.items() if v[1]}


def main():
    dev = torch.device('cuda', 0)
    out = {}
    # ---- config-2 tile: standalone functions + RLE on its panoptic map
    H = W = 4096
    d = synth_tile(H, W, 500, seed=0)
    sem, hm, off = (torch.from_numpy(d[k]).to(dev) for k in ('sem', 'ctr_hmp', 'offsets'))
    pan, ctr = pp.get_panoptic_segmentation(sem, hm, off, [1], 1000, 64, 0, 0.1, 7)
    npx = H * W
    out['tile'] = {'shape': [H, W], 'K': int(ctr.shape[1])}
    ms = timed(lambda: pp.find_instance_center(hm, 0.1, 7))
    out['find_instance_center'] = {'ms_incl_host_readback': ms, 'stages': stages(lambda: pp.find_instance_center(hm, 0.1, 7))}
    ms = timed(lambda: pp.group_pixels(ctr[0], off))
    out['group_pixels_all_pixels'] = {'ms': ms, 'GBps_alg16': 16 * npx / ms / 1e6, 'stages': stages(lambda: pp.group_pixels(ctr[0], off))}
    ins, _ = pp.get_instance_segmentation(sem[0], hm, off, [1], 0.1, 7)
    ms = timed(lambda: pp.merge_semantic_and_instance(sem[0], ins, 1000, [1], 64, 0))
    out['merge_semantic_and_instance'] = {'ms_incl_aminmax_readback': ms,
                                          'stages': stages(lambda: pp.merge_semantic_and_instance(sem[0], ins, 1000, [1], 64, 0))}
    p2 = pan[0, 0].contiguous()
    st = stages(lambda: rle.rle_tables(p2, [1], 1000, [1], True))
    ms = timed(lambda: rle.rle_tables(p2, [1], 1000, [1], True), n=10)
    inst, runs = rle.rle_tables(p2, [1], 1000, [1], True)
    out['pan_seg_to_rle (4096^2)'] = {'ms_incl_readback_and_d2h': ms, 'stages': st, 'instances': int(inst.shape[0]), 'runs': int(runs.shape[0]),
                                      'rle_mark_GBps_alg8': 8 * npx / st.get('rle_mark', float('nan')) / 1e6}
    # ---- config-3 slice: median(3)+harden on 2048^2 probabilities, coarse path
    S = 2048
    sl = [s for i, s in enumerate(synth_stack_slices(44, S, S, 400, seed=5)) if i >= 41]
    planes = [torch.from_numpy(s['sem_prob']).to(dev) for s in sl]
    chm, coff = torch.from_numpy(sl[1]['ctr_hmp']).to(dev), torch.from_numpy(sl[1]['offsets']).to(dev)
    st = stages(lambda: engines.median_harden(planes, 0.3, want_median=True, want_sem='u8'))
    out['median3_harden (2048^2)'] = {'stages': st, 'GBps_alg(3x4 in + 4 + 1 out)': 17 * S * S / st['median_harden'] / 1e6}
    eng = engines.PanopticDeepLabRenderEngine(torch.nn.Identity(), thing_list=[1], label_divisor=20000, stuff_area=64,
                                              void_label=0, nms_threshold=0.1, nms_kernel=3, confidence_thr=0.3)
    st = stages(lambda: eng._fused_postprocess(planes[1], chm, coff, 1))
    ms = timed(lambda: eng._fused_postprocess(planes[1], chm, coff, 1))
    out['render_fused_postprocess (2048^2, coarse 512^2)'] = {'ms_incl_host_readbacks': ms, 'stages': st, 'device_ms': round(sum(st.values()), 4)}
    pan3 = eng._fused_postprocess(planes[1], chm, coff, 1)
    st = stages(lambda: rle.rle_tables(pan3[0], [1], 20000, [1], True))
    ms = timed(lambda: rle.rle_tables(pan3[0], [1], 20000, [1], True), n=10)
    out['pan_seg_to_rle (2048^2)'] = {'ms_incl_readback_and_d2h': ms, 'stages': st, 'device_ms': round(sum(st.values()), 4)}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
