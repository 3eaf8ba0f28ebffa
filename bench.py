#!/usr/bin/env python
"""
bench.py — panoptic post-processing throughput on B200 (BASELINE.json configs[1]).

One "step" = one pass of the fused get_panoptic_segmentation pipeline over a batch of 16
synthetic 4096x4096 tiles (~500 centers per tile) per GPU.

  value     Mpix/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same metric through the host-buffer C-ABI entry (emp_panoptic_batched_host): pinned
            host tensors in, H2D + kernels + D2H of the int64 panoptic maps inside the timed region
  roofline  the dominant kernel (assign: sem + offsets -> code map) against the measured HBM copy peak
            (MEASURED_PEAKS.json); traffic = ncu DRAM bytes per launch from profiles/traffic.json
  cpu_baseline  the reference's CPU op sequence (oracle/torch_port.py) on a bounded crop, rank 0, N=1

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 4096
TILES = 16
N_INST = 500
THINGS = [1]
LABEL_DIVISOR, STUFF_AREA, VOID, THR, NMS_K = 1000, 64, 0, 0.1, 7
ALG_BYTES_PER_PX = 28          # sem i64 8 + heat-map f32 4 + offsets 2 x f32 8 in, pan i64 8 out
# each kernel of the chain owns one of the four algorithmic streams
STAGE_ALG_BYTES_PER_PX = {'nms_peaks': 4, 'assign': 16, 'apply_lut': 8}
METRIC, UNIT = 'panoptic_postproc_throughput', 'Mpix/s'
WORKLOAD = 'postproc_16x4096x4096_k500'


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_JSON_FD = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: libraries that print there (NCCL's version banner on the first
    communicator, nvidia-smi children) are pointed at stderr instead."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def bind_cpu_slice(local_rank, world):
    """Give this rank its own slice of the host's CPUs (and so, by first touch, of its memory): N ranks that all run on
    the same cores fight over them in the host-buffer path (round 1: 0.21 e2e efficiency at 8 ranks).  The library sizes
    its worker pool from the affinity mask."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // max(world, 1))
        mine = cpus[local_rank * per:(local_rank + 1) * per] or cpus
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception as e:                                   # best effort: an unbound rank is still correct
        return f'unbound ({type(e).__name__})'


def ncu_traffic(kernel):
    """DRAM bytes per pixel of `kernel` from the committed ncu capture of this command
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch / pixels per launch)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return float(json.load(f)['dram_bytes_per_px'][kernel])
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def mark(self):
        return len(self.rows)

    def stop(self, start=0):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for r in self.rows[start:] if len(r) >= 6] or [r for r in self.rows if len(r) >= 6]
        if not rows:
            return None
        sm = [float(r[0]) for r in rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(rows)}


def reference_postprocess():
    """The reference's own post-processing module when build() staged it (oracle/_ref/empanada_inference_postprocess.py: the unmodified
    file, copied from /root/reference in the build container; git-ignored, it travels with the snapshot), else the
    oracle's op-sequence port.  Returns (find_instance_center, group_pixels, merge_semantic_and_instance, kind)."""
    path = os.path.join(ROOT, 'oracle', '_ref', 'empanada_inference_postprocess.py')
    if os.path.exists(path):
        import importlib.util
        spec = importlib.util.spec_from_file_location('empanada_reference_postprocess', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.find_instance_center, mod.group_pixels, mod.merge_semantic_and_instance, 'reference'
    from oracle import torch_port as tp
    return (lambda hm, thr, k: tp.centers_by_maxpool(hm, thr, k), lambda ctr, off: tp.pixel_ids(ctr, off),
            tp.vote_and_paste, 'port')


_REF_TILE = {}


def cpu_reference_sample(seed=0, win_centers=1024, win_group=160, win_merge=512):
    """One bounded sample of the workload on the host cores, through the reference's own functions
    (postprocess.py:38-76, :118-169, :223-296), with the per-pixel cost of the FULL tile:

      find_instance_center  on a win_centers^2 window of the tile's heat-map            (x 4096^2 / win_centers^2)
      group_pixels          on a win_group^2 window of the offsets with ALL ~500 centers of the tile, i.e. the tile's
                            25 chunks of 20 centers per pixel                            (x 4096^2 / win_group^2)
      merge_semantic_and_instance  on a win_merge^2 window (its instance loop costs instances x pixels)
                                                                  (x 4096^2 / win_merge^2 x K / instances in the window)
    Returns (Mpix/s of a whole tile extrapolated from the three timings, seconds spent, K, threads, description)."""
    import torch
    from empanada_b200.synth import synth_tile
    find_centers, group_pixels, merge, kind = reference_postprocess()
    if seed not in _REF_TILE:
        _REF_TILE.clear()
        _REF_TILE[seed] = synth_tile(H, W, N_INST, seed)
    d = _REF_TILE[seed]
    sem, hm, off = (torch.from_numpy(d[k]) for k in ('sem', 'ctr_hmp', 'offsets'))
    # all centers of the tile (not timed: the oracle's closed form; the timed part below runs the reference)
    import oracle
    ctr_all = torch.from_numpy(oracle.find_instance_center(d['ctr_hmp'], THR, NMS_K))
    K = int(ctr_all.shape[0])
    y0 = x0 = (H - win_centers) // 2
    t0 = time.perf_counter()
    find_centers(hm[:, :, y0:y0 + win_centers, x0:x0 + win_centers].contiguous(), THR, NMS_K)
    t_c = time.perf_counter() - t0
    # a window around a center, centers shifted into its frame: every pixel still meets all K centers
    cy, cx = [int(v) for v in ctr_all[K // 2]]
    gy, gx = min(max(cy - win_group // 2, 0), H - win_group), min(max(cx - win_group // 2, 0), W - win_group)
    shifted = ctr_all - torch.tensor([gy, gx])
    t0 = time.perf_counter()
    ids = group_pixels(shifted, off[:, :, gy:gy + win_group, gx:gx + win_group].contiguous())
    t_g = time.perf_counter() - t0
    my, mx = min(max(cy - win_merge // 2, 0), H - win_merge), min(max(cx - win_merge // 2, 0), W - win_merge)
    sem_w = sem[:, :, my:my + win_merge, mx:mx + win_merge].contiguous()
    ins_w = torch.from_numpy(d['ins'][my:my + win_merge, mx:mx + win_merge].astype('int64'))[None] * (sem_w[0] > 0)   # GT ids: same instance count / sizes
    k_w = max(int(torch.unique(ins_w).numel()) - 1, 1)
    t0 = time.perf_counter()
    merge(sem_w, ins_w, LABEL_DIVISOR, THINGS, STUFF_AREA, VOID)
    t_m = time.perf_counter() - t0
    n_px = H * W
    t_tile = t_c * n_px / win_centers ** 2 + t_g * n_px / win_group ** 2 + t_m * (n_px / win_merge ** 2) * (K / k_w)
    desc = (f'{kind}: find_instance_center on {win_centers}^2 + group_pixels on {win_group}^2 with all {K} centers of the tile '
            f'(25 chunks) + merge on {win_merge}^2 ({k_w} instances), each scaled to the 4096^2 tile '
            f'({t_c:.2f} / {t_g:.2f} / {t_m:.2f} s measured -> {t_tile:.0f} s per tile); torch {torch.__version__} CPU')
    return n_px / t_tile / 1e6, t_c + t_g + t_m, K, torch.get_num_threads(), desc, kind


def run_config1(dev, reps=50):
    """BASELINE configs[0]: the 1024 x 1024 tile the reference itself was run on (tests/golden/make_config1.py: its
    ResNet-50 PanopticDeepLab + engine on the CPU) through this package's PanopticDeepLabEngine (engines.py:92-160 of the
    reference: sigmoid, harden, get_panoptic_segmentation), the CNN replaced by its recorded outputs.  Latency per tile
    as a caller sees it (Python API, synchronised), checked against the reference's panoptic map."""
    import torch
    from empanada_b200.inference import engines as eng
    from empanada_b200.synth import CONFIG1, CONFIG1_TILE, config1_heads
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'config1.npz'))
    heads = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in config1_heads(g['consts']).items()}

    class RecordedHeads(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.anchor = torch.nn.Parameter(torch.zeros(1, device=dev))       # the engine asks the model for its device

        def forward(self, x):
            return dict(heads)

    engine = eng.PanopticDeepLabEngine(RecordedHeads(), **CONFIG1)
    hw = CONFIG1_TILE[0]
    image = torch.zeros((1, 1, hw, hw), device=dev)
    for _ in range(5):
        pan = engine(image)
    torch.cuda.synchronize(dev)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        pan = engine(image)
        torch.cuda.synchronize(dev)
        times.append(time.perf_counter() - t0)
    ms = statistics.median(times) * 1e3
    rec = {'workload': f'postproc_1x{hw}x{hw}_k{CONFIG1_TILE[1]}_engine2d', 'ms_per_tile': ms, 'value': hw * hw / (ms * 1e-3) / 1e6, 'unit': UNIT,
           'what': 'PanopticDeepLabEngine.__call__ minus the CNN: sigmoid + harden + fused post-processing, one synchronised Python call per tile',
           'matches_reference_run': bool(np.array_equal(pan.cpu().numpy().reshape(hw, hw), g['pan'].reshape(hw, hw)))}
    ref = os.path.join(ROOT, 'profiles', 'r2_config1_cpu.json')
    if os.path.exists(ref):
        with open(ref) as f:
            r = json.load(f)
        rec['reference_cpu_build_container'] = {k: r[k] for k in ('threads', 'seconds_cnn_forward', 'seconds_postprocess', 'postprocess_mpix_per_s')}
    return rec


def workload_config(world, B, Ks, dense=False):
    return {'workload': WORKLOAD if not dense else 'postproc_16x4096x4096_dense_k5000', 'tiles_per_gpu': B, 'tile': [H, W],
            'centers_per_tile': Ks, 'thing_list': THINGS, 'label_divisor': LABEL_DIVISOR, 'nms_kernel': NMS_K,
            'l2': f'inputs {B * H * W * 20 / 1e9:.1f} GB per step >> 126 MB L2, no flush needed',
            'parallelism': f'dp{world} (independent tiles per rank, no collective on the data path)'}


def run_reference(args, rank):
    """--impl reference: the reference's CPU post-processing on the host cores, all threads, one bounded sample of the
    workload per step (cpu_reference_sample: full-tile per-pixel cost, extrapolated to Mpix/s of whole tiles)."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vals, sample, kind = [], '', 'port'
    for i in range(args.warmup + args.steps):
        v, dt, k, threads, sample, kind = cpu_reference_sample(seed=0)
        log(f'[reference] step {i}: {v:.5f} Mpix/s ({dt:.1f} s of samples, K={k})')
        if i >= args.warmup:
            vals.append((v, dt))
    value = len(vals) / sum(1.0 / v for v, _ in vals)              # tiles per second over the timed steps, in Mpix/s
    ms = 1e3 * sum(dt for _, dt in vals) / len(vals)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int64/f32', 'data': 'synthetic',
            'config': workload_config(args.gpus, TILES, [N_INST - 2, N_INST]),
            'sample': sample,
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': kind,
                             'sample': sample + '; one sample per step, ms_per_step = seconds of samples per step'},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--tiles', type=int, default=TILES)
    ap.add_argument('--instances', type=int, default=N_INST)
    ap.add_argument('--thr', type=float, default=THR, help='center threshold (experiments; the workload uses 0.1)')
    ap.add_argument('--dense', action='store_true', help='BASELINE configs[4] as the main workload: ~5000 instances of semi-axes 4..12 px per tile')
    ap.add_argument('--no-stack', action='store_true', help='skip the configs[2] stack sub-record')
    ap.add_argument('--no-deep', action='store_true', help='skip the 2048-slice variant of the stack sub-record')
    ap.add_argument('--no-cnn', action='store_true', help='skip the stack loop with the stand-in CNN in it')
    ap.add_argument('--no-dense', action='store_true', help='skip the configs[4] dense sub-record (N = 1)')
    ap.add_argument('--dense-tiles', type=int, default=8)
    ap.add_argument('--no-config1', action='store_true', help='skip the configs[0] sub-record (one 1024^2 tile through the 2D engine, N = 1)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    claim_stdout()

    if args.impl == 'reference':
        run_reference(args, rank)               # a few seconds of CPU per step: --steps 20 --warmup 5 ends within minutes
        return

    import torch
    import torch.distributed as dist
    from empanada_b200 import _cabi as C
    from empanada_b200.inference import postprocess as pp
    from empanada_b200.synth import synth_tile
    import ctypes

    args.warmup = max(args.warmup, 3)          # timing rules: at least 3 warm-up steps
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        log(f'[rank {rank}] host threads bound to a slice of {bind_cpu_slice(local_rank, world)} cpus')
    B = args.tiles
    n_px = H * W

    # ---- synthetic batch: pinned host tensors (the e2e path reads them) + resident device copies
    t0 = time.time()
    sem_h = torch.empty((B, H, W), dtype=torch.int64).pin_memory()
    hm_h = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
    off_h = torch.empty((B, 2, H, W), dtype=torch.float32).pin_memory()
    pan_h = torch.empty((B, H, W), dtype=torch.int64).pin_memory()
    for b in range(B):
        if args.dense:
            d = synth_tile(H, W, 5000, seed=rank * 1000 + b, semi_axes=(4.0, 12.0), sigma=2.0)
        else:
            d = synth_tile(H, W, args.instances, seed=rank * 1000 + b)
        sem_h[b] = torch.from_numpy(d['sem'][0, 0])
        hm_h[b] = torch.from_numpy(d['ctr_hmp'][0, 0])
        off_h[b] = torch.from_numpy(d['offsets'][0])
    sem, hm, off = sem_h.to(dev), hm_h.to(dev), off_h.to(dev)
    log(f'[rank {rank}] synthetic batch ready in {time.time() - t0:.1f} s')

    L = C.lib()
    things, nt = C.i64_array(THINGS)
    k_cap = pp.DEFAULT_K_CAP
    per_tile = L.emp_workspace_bytes(H, W, k_cap, nt)
    ws = torch.empty(per_tile * B, dtype=torch.uint8, device=dev)
    pan = torch.empty((B, H, W), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        C.check(L.emp_panoptic_batched(B, sem.data_ptr(), 0, hm.data_ptr(), off.data_ptr(), H, W, things, nt,
                                       LABEL_DIVISOR, STUFF_AREA, VOID, args.thr, NMS_K, pan.data_ptr(), None, 0, k_cap,
                                       ws.data_ptr(), per_tile, ctypes.c_void_p(stream.cuda_stream)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    st = C.read_status(ws, B, per_tile).reshape(B, -1)
    Ks = [int(v) for v in st[:, C.ST_K]]
    assert all(int(f) == 0 for f in st[:, C.ST_FLAGS]), 'status flags set'
    log(f'[rank {rank}] centers per tile: min {min(Ks)} max {max(Ks)}')

    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.5 if sampler else 0)
    mark = sampler.mark() if sampler else 0
    C.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = C.profile_read()
    C.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B * n_px / (ms_step * 1e-3) / 1e6

    # ---- end to end: pinned host buffers through emp_panoptic_batched_host ---------------------
    scratch_bytes = L.emp_host_scratch_bytes(H, W, k_cap, nt)
    scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
    k_out = (ctypes.c_int32 * B)()
    f_out = (ctypes.c_int32 * B)()

    def e2e_step():
        C.check(L.emp_panoptic_batched_host(B, sem_h.data_ptr(), hm_h.data_ptr(), off_h.data_ptr(), H, W, things, nt,
                                            LABEL_DIVISOR, STUFF_AREA, VOID, args.thr, NMS_K, pan_h.data_ptr(), k_out, f_out,
                                            k_cap, scratch.data_ptr(), scratch_bytes))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()          # blocks until the last D2H copy has landed
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B * n_px / (e2e_ms * 1e-3) / 1e6
    assert list(k_out) == Ks and all(f == 0 for f in f_out)
    sem_bpp = float(L.emp_host_sem_bytes_per_px())     # 1.0 when every tile's class map was narrowed on the host
    same = all(bool(torch.equal(pan_h[b], pan[b].cpu())) for b in range(B))      # every tile, not just the last
    # the same bytes over the same link with no kernel in between: what the host fabric allows this rank right now
    sem8_h = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
    sem8_d = torch.empty((H, W), dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copy_floor():
        for b in range(B):
            with torch.cuda.stream(s_in):
                sem8_d.copy_(sem8_h[b], non_blocking=True)
                hm[b].copy_(hm_h[b], non_blocking=True)
                off[b].copy_(off_h[b], non_blocking=True)
            with torch.cuda.stream(s_out):
                pan_h[b].copy_(pan[b], non_blocking=True)
        torch.cuda.synchronize(dev)

    copy_floor()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        copy_floor()
    floor_ms = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
    if world > 1:
        t = torch.tensor([floor_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        floor_ms = float(t.item())
    del sem8_h
    clocks = sampler.stop(mark) if sampler else None

    # ---- BASELINE configs[2]: the z-sharded stack (strong scaling: the 512 x 2048^2 volume is split over the ranks) ----
    stack_rec = None
    if not args.no_stack:
        import bench_stack as bs
        t0 = time.time()
        slices = bs.make_slices(dev, 2048)
        log(f'[rank {rank}] stack: {len(slices)} distinct slices ready in {time.time() - t0:.1f} s')
        stack_rec, s_out_, _, _ = bs.run_stack(dev, rank, world, slices, 512, 2048, ks=3, repeats=5, warmup=4, profile=(world == 1))
        if world > 1:
            bs.add_parity(stack_rec, s_out_, dev, rank, world, slices, 512, 2048, 3, 0, 4096)
        peak_, _src = measured_peak()
        stack_rec['roofline'] = {'bound': 'hbm', 'alg_bytes_per_voxel': bs.ALG_BYTES_PER_VOXEL,
                                 'achieved': bs.ALG_BYTES_PER_VOXEL * stack_rec['value'] / 1e9 / world, 'peak': peak_, 'unit': 'GB/s per GPU',
                                 'frac': bs.ALG_BYTES_PER_VOXEL * stack_rec['value'] / 1e9 / world / peak_,
                                 'note': 'host-visible time (enqueue, kernels, D2H of the tables, collectives) against the algorithmic bytes of '
                                         'SURVEY 8d; the fused path never writes or re-reads the int64 label map those bytes include'}
        del s_out_
        if not args.no_deep:
            # the same path on a 4 x deeper stack: per-rank work large enough that the fixed cost of a block
            # (a few tenths of a millisecond of launches, carry exchange and table read-back) stops dominating at 8 ranks
            deep, d_out, _, _ = bs.run_stack(dev, rank, world, slices, 2048, 2048, ks=3, repeats=3, warmup=2)
            if world > 1:
                bs.add_parity(deep, d_out, dev, rank, world, slices, 2048, 2048, 3, 0, 4096)
            stack_rec['deep'] = {k: deep.get(k) for k in ('value', 'unit', 'seconds', 'seconds_all', 'parity', 'single_gpu_seconds_same_run',
                                                         'speedup_vs_n1', 'config')}
            del d_out
        if not args.no_cnn:
            stack_rec['with_cnn'] = bs.run_stack_with_cnn(dev, rank, world, slices, 512, 2048, 3)
        del slices

    # ---- BASELINE configs[4]: dense-instance stress, a few tiles through the same fused entry point ----
    dense_rec = None
    if world == 1 and not args.no_dense and not args.dense:
        t0 = time.time()
        Bd = args.dense_tiles
        ds = [synth_tile(H, W, 5000, seed=7000 + b, semi_axes=(4.0, 12.0), sigma=2.0) for b in range(Bd)]
        sem_d = torch.stack([torch.from_numpy(d['sem'][0, 0]) for d in ds]).to(dev)
        hm_d = torch.stack([torch.from_numpy(d['ctr_hmp'][0, 0]) for d in ds]).to(dev)
        off_d = torch.stack([torch.from_numpy(d['offsets'][0]) for d in ds]).to(dev)
        del ds
        log(f'[rank 0] dense tiles ready in {time.time() - t0:.1f} s')

        def dense_step():
            C.check(L.emp_panoptic_batched(Bd, sem_d.data_ptr(), 0, hm_d.data_ptr(), off_d.data_ptr(), H, W, things, nt,
                                           LABEL_DIVISOR, STUFF_AREA, VOID, THR, NMS_K, pan.data_ptr(), None, 0, k_cap,
                                           ws.data_ptr(), per_tile, ctypes.c_void_p(stream.cuda_stream)))
        for _ in range(3):
            dense_step()
        torch.cuda.synchronize(dev)
        Kd = [int(v) for v in C.read_status(ws, Bd, per_tile).reshape(Bd, -1)[:, C.ST_K]]
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(stream)
        for _ in range(10):
            dense_step()
        d1.record(stream)
        torch.cuda.synchronize(dev)
        dms = d0.elapsed_time(d1) / 10
        dense_rec = {'workload': f'postproc_{Bd}x4096x4096_dense_k5000', 'value': Bd * n_px / (dms * 1e-3) / 1e6, 'unit': UNIT,
                     'ms_per_step': dms, 'tiles': Bd, 'centers_per_tile': [min(Kd), max(Kd)],
                     'hbm_frac_of_measured': ALG_BYTES_PER_PX * Bd * n_px / (dms * 1e-3) / 1e9 / measured_peak()[0]}
        del sem_d, hm_d, off_d

    # ---- BASELINE configs[0]: the reference's CPU-runnable case — one 1024^2 tile through PanopticDeepLabEngine ----
    config1_rec = None
    if world == 1 and not args.no_config1:
        config1_rec = run_config1(dev)

    if rank == 0:
        peak, peak_src = measured_peak()
        dom = max(STAGE_ALG_BYTES_PER_PX, key=lambda k: prof[k][0])          # the kernel that takes the most time
        a_ms, a_n = prof[dom]
        launches_per_step = a_n / args.steps
        px_per_launch = B * n_px / launches_per_step
        dom_ms = a_ms / max(a_n, 1)
        achieved = STAGE_ALG_BYTES_PER_PX[dom] * px_per_launch / (dom_ms * 1e-3) / 1e9
        stage_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1]}
        launches = sum(v[1] for k, v in prof.items() if k != 'memset')          # the clears are not kernels of this library
        pipeline_gbs = ALG_BYTES_PER_PX * B * n_px / (ms_step * 1e-3) / 1e9
        tpp = ncu_traffic(dom)
        traffic = tpp * px_per_launch if tpp is not None else None
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'int64/f32', 'data': 'synthetic',
            'config': workload_config(world, B, [min(Ks), max(Ks)], args.dense),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(B * n_px * (12 + sem_bpp)), 'd2h_bytes_per_step': B * n_px * 8 + B * 64,
                    'host_buffer_bytes_per_step': B * n_px * 20, 'sem_bytes_per_px_on_the_link': sem_bpp,
                    'ms_per_step': e2e_ms, 'steps': args.e2e_steps, 'matches_resident_result': same,
                    'copy_only_floor_ms_per_step': floor_ms,
                    'copy_only_floor_note': 'the same H2D / D2H bytes on two streams with no kernel launched, max over ranks: the '
                                            'share of e2e time the host memory / PCIe fabric of the box dictates',
                    'api': 'emp_panoptic_batched_host (pinned host tensors, 3-slot H2D/compute/D2H pipeline; int64 class maps '
                           'narrowed to uint8 by host worker threads before the link)'},
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'alg_bytes_per_px': STAGE_ALG_BYTES_PER_PX[dom], 'avg_launch_ms': dom_ms,
                         'note': ('achieved counts the algorithmic bytes of the stage (SURVEY 8d: sem 8 + offsets 8 B/px for assign); '
                                  'traffic is what ncu saw in DRAM for one launch — below it, because offsets are fetched only for '
                                  'strips holding thing pixels and background strips store no codes'),
                         'pipeline': {'alg_bytes_per_px': ALG_BYTES_PER_PX, 'achieved': pipeline_gbs,
                                      'frac': pipeline_gbs / peak, 'frac_of_8TBs_spec': pipeline_gbs / 8000.0},
                         'stage_ms_per_step': stage_ms},
            'clocks': clocks,
            'stack': stack_rec,
            'dense': dense_rec,
            'config1': config1_rec,
        }
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            v, dt, k, threads, sample, kind = cpu_reference_sample()
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': kind,
                                    'sample': f'{sample}; {dt:.1f} s of CPU'}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
