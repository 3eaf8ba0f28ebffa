"""
Z-sharded stack inference: the part of the reference's 3D path between "CNN heads produced
sem / ctr_hmp / offsets for slice z" and "the host tracker receives slice z's RLE dict"
(reference scripts/pdl_inference3d.py:163-176 + empanada/inference/patterns.py:68-100, and the
multi-GPU recipe patterns.forward_multigpu :279-350), one process per GPU.

Sharding.  Slices are independent except for the recursive median queue (engines.py:47-90):
    m_i = median(f_{i-mid}, ..., f_{i-1}, s_i, s_{i+1}, ..., s_{i+mid}),   f_j = m_j (j >= mid) else s_j
for mid <= i < D - mid, raw s_i otherwise.  Rank r owns the contiguous block [z0, z1).  It needs
  * a look-ahead halo: the raw probabilities of slices z1 .. z1+mid-1 (it runs the CNN on them too);
  * a carry: the filtered planes f_{z0-mid} .. f_{z0-1} from rank r-1 (one send/recv of `mid`
    (C,H,W) fp32 planes over NCCL/NVLink), available once r-1 has run its — elementwise, cheap —
    median chain.  The CNN forwards, which dominate, never wait on it.
With median_kernel_size == 1 nothing is exchanged.  With median_kernel_size == 3 (the scripts' default) no rank waits
for the chain of the rank below: a median of three is a clamp of one argument to the range of the other two, clamps
compose, so every rank first reduces its block to ONE clamp from its own raw planes (emp_median3_compose), the carry
crosses the ranks with one clamp per rank, and all chains then run at once (exchange_carry_median3).

Labels.  Every slice numbers its instances 1..n per class.  So that labels from different ranks
never collide before the host-side cross-slice matcher renumbers them, ranks all-gather their
per-class maximum instance count (one int64 per class) and add the exclusive prefix as an offset;
rank 0 keeps offset 0, which is what the reference's matcher sees for the first slice.
"""
import contextlib
import ctypes
import gc
import math
import time

import numpy as np
import torch
import torch.distributed as dist

__all__ = ['partition_slices', 'halo_range', 'median_chain', 'exchange_carry', 'exchange_carry_median3', 'compose_median3',
           'label_offsets',
           'apply_label_offset', 'StackShard']


@contextlib.contextmanager
def _gc_paused():
    """The block loops below build ~10^5 small dicts that are all kept; CPython's cyclic collector re-scans the
    growing pile every few thousand allocations (measured: 0.30 s instead of 0.13 s to assemble a 512-slice block) and
    none of it can be garbage, so collection is paused for the duration."""
    was_on = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_on:
            gc.enable()


_stream_cache = {}


def _side_streams(device, n):
    """n side streams of `device`, created once per process: the per-stream scratch buffers (C.workspace) and the
    caching allocator's per-stream pools stay warm from block to block."""
    have = _stream_cache.setdefault(device.index, [])
    while len(have) < n:
        have.append(torch.cuda.Stream(device))
    return have[:n]


def partition_slices(depth, world_size, rank):
    """Contiguous z-block [z0, z1) of `rank`: the first depth % world_size ranks get one extra slice."""
    base, extra = divmod(depth, world_size)
    z0 = rank * base + min(rank, extra)
    return z0, z0 + base + (1 if rank < extra else 0)


def halo_range(depth, world_size, rank, median_kernel_size):
    """Slices whose head tensors `rank` must compute: its block plus the look-ahead halo."""
    z0, z1 = partition_slices(depth, world_size, rank)
    mid = (median_kernel_size - 1) // 2
    return z0, min(depth, z1 + mid)


def median_chain(raw, z0, z1, depth, ks, carry, median_fn):
    """Recursive median over the block [z0, z1).

    raw      dict {z: sem plane} for z in [z0, min(depth, z1 + mid))
    carry    list of the `mid` planes f_{z0-mid} .. f_{z0-1} (filtered where z >= mid, else raw);
             ignored for z0 == 0
    median_fn(list of ks planes) -> plane   (libempanada_b200's emp_median_harden on the GPU,
             any odd-count median on the CPU in tests)
    Returns (filtered {z: plane} for z in [z0, z1), carry for the next rank).
    """
    mid = (ks - 1) // 2
    assert depth >= ks, 'stack shallower than the median kernel'
    f = {}
    if z0 > 0:
        assert len(carry) == mid
        for j, p in enumerate(carry):
            f[z0 - mid + j] = p
    for z in range(z0, z1):
        if z < mid or z >= depth - mid or ks == 1:
            f[z] = raw[z]                                   # queue still filling / end(): raw
        else:
            window = [f[j] for j in range(z - mid, z)] + [raw[j] for j in range(z, z + mid + 1)]
            f[z] = median_fn(window)
    out = {z: f[z] for z in range(z0, z1)}
    nxt = [f[z] for z in range(z1 - mid, z1)] if mid > 0 else []
    return out, nxt


def exchange_carry(chain_fn, rank, world_size, mid, plane_like, group=None):
    """Run `chain_fn(carry) -> (result, next_carry)` on every rank in rank order, handing the
    carry planes from rank r to r+1 with point-to-point send/recv (NCCL on CUDA tensors, gloo on
    CPU tensors).  plane_like: a tensor with the carry planes' shape / dtype / device."""
    carry = []
    if mid > 0 and rank > 0 and world_size > 1:
        carry = [torch.empty_like(plane_like) for _ in range(mid)]
        for p in carry:
            dist.recv(p, src=rank - 1, group=group)
    result, nxt = chain_fn(carry)
    if mid > 0 and rank + 1 < world_size:
        for p in nxt:
            dist.send(p.contiguous(), dst=rank + 1, group=group)
    return result


def compose_median3(raw, z0, z1, depth, like=None):
    """The block [z0, z1) of a median_kernel_size == 3 chain as ONE clamp: returns (A, B) with
    f_{z1-1} = min(max(f_{z0-1}, A), B).  A median of three is a clamp of one argument to the range of the other
    two, f_i = clamp(f_{i-1}; min(s_i, s_{i+1}), max(s_i, s_{i+1})), clamps compose into clamps, and the raw first /
    last slice of the stack is the constant clamp (s_i, s_i).  torch restatement of libempanada_b200's
    emp_median3_compose for tensors on any device (StackShard uses the kernel; the gloo tests use this)."""
    ref = raw[z0] if z0 < z1 else like
    A = torch.full_like(ref, float('-inf'))
    B = torch.full_like(ref, float('inf'))
    for z in range(z0, z1):
        if z < 1 or z >= depth - 1:
            lo = hi = raw[z]
        else:
            lo, hi = torch.minimum(raw[z], raw[z + 1]), torch.maximum(raw[z], raw[z + 1])
        A = torch.minimum(torch.maximum(A, lo), hi)
        B = torch.minimum(torch.maximum(B, lo), hi)
    return A, B


def exchange_carry_median3(compose_fn, chain_fn, rank, world_size, plane_like, group=None):
    """exchange_carry for median_kernel_size == 3 without the rank-after-rank wait: every rank first reduces its
    block to one clamp (compose_fn() -> (A, B), from its own raw planes only), the carry plane then crosses the ranks
    with ONE clamp per rank, and the real chains (chain_fn(carry) -> (result, next_carry)) run on all ranks at once.
    Returns (result, mismatch): mismatch is a 0-d int64 tensor, non-zero when the plane this rank handed on is not
    the chain's own last plane (only possible with NaNs) — the caller must then redo the block with exchange_carry."""
    carry, sent = [], None
    if rank + 1 < world_size:
        A, B = compose_fn()
    if rank > 0:
        carry = [torch.empty_like(plane_like)]
        dist.recv(carry[0], src=rank - 1, group=group)
    if rank + 1 < world_size:
        sent = torch.minimum(torch.maximum(carry[0], A), B) if rank > 0 else torch.minimum(A, B)
        dist.send(sent.contiguous(), dst=rank + 1, group=group)
    result, nxt = chain_fn(carry)
    if sent is None:
        return result, torch.zeros((), dtype=torch.int64, device=plane_like.device)
    own = nxt[0] if nxt else sent                       # an empty block hands the plane through
    return result, (own != sent).any().to(torch.int64)  # NaN != NaN: flagged, which is what we want


def label_offsets(max_counts, group=None):
    """max_counts: int64 tensor (n_classes,) — this rank's largest per-slice instance count per
    class.  All-gathers it (NCCL over NVLink for CUDA tensors) and returns this rank's exclusive
    prefix (n_classes,) plus the gathered (world, n_classes) table."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return torch.zeros_like(max_counts), max_counts[None].clone()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = [torch.zeros_like(max_counts) for _ in range(world)]
    dist.all_gather(gathered, max_counts, group=group)
    table = torch.stack(gathered)
    return table[:rank].sum(0), table


def apply_label_offset(rle_seg, offsets_by_class, label_divisor, thing_list):
    """Shift the instance labels of thing classes in one slice's rle dict by this rank's offset."""
    out = {}
    for cls, attrs in rle_seg.items():
        off = int(offsets_by_class.get(cls, 0)) if cls in thing_list else 0
        if off == 0:
            out[cls] = attrs
            continue
        shifted = {}
        for lab, a in attrs.items():
            new = lab + off
            if new >= (cls + 1) * label_divisor:
                raise ValueError(f'label {lab} + offset {off} leaves class {cls}\'s range; '
                                 f'raise label_divisor (reference default for 3D is 20000)')
            shifted[new] = a
        out[cls] = shifted
    return out


def _slice_sync(engine, head, sem_prob, labels, upsampling, force_connected):
    """One slice, synchronously (status read-backs and retries inside): the fallback for a slice
    whose deferred run overflowed a table."""
    from empanada_b200.inference import rle
    pan = engine._fused_postprocess(sem_prob, head['ctr_hmp'], head['offsets'], upsampling)
    if head['size'] is not None:
        pan = pan[..., :head['size'][0], :head['size'][1]]
    return rle.pan_seg_to_rle_seg(pan, labels, engine.label_divisor, engine.thing_list, force_connected)


class StackShard:
    """One rank's share of a stack: feed it the head tensors of its slices (block + halo) in z
    order, then `finish()` returns {z: rle_seg} for the block.

    engine: a PanopticDeepLabRenderEngine(3d) (its post-processing parameters and fused kernels
    are used; its own median queue is bypassed in favour of the sharded chain above).
    """

    def __init__(self, engine, labels, depth, rank=0, world_size=1, median_kernel_size=3,
                 upsampling=1, force_connected=True, group=None, n_streams=4):
        assert median_kernel_size % 2 == 1, "Kernel size must be odd integer!"
        assert math.log(upsampling, 2).is_integer(), "Upsampling factor not log base 2!"
        self.engine, self.labels, self.depth = engine, list(labels), depth
        self.rank, self.world, self.ks = rank, world_size, median_kernel_size
        self.mid = (median_kernel_size - 1) // 2
        self.upsampling, self.force_connected, self.group = upsampling, force_connected, group
        self.z0, self.z1 = partition_slices(depth, world_size, rank)
        _, self.z_halo = halo_range(depth, world_size, rank, median_kernel_size)
        self.heads = {}
        self.n_streams = max(1, int(n_streams))   # side streams the block's slices are spread over
        self._streams = None

    def slices(self):
        """z indices this rank must run the CNN on, in order."""
        return range(self.z0, self.z_halo)

    def add(self, z, sem_prob, ctr_hmp=None, offsets=None, size=None):
        """Head tensors of slice z: sem_prob (1,C,H,W) probabilities; ctr_hmp/offsets only needed
        for z inside the block (halo slices contribute their probabilities only)."""
        assert self.z0 <= z < self.z_halo
        self.heads[z] = {'sem': sem_prob, 'ctr_hmp': ctr_hmp, 'offsets': offsets, 'size': size}

    def _finish_block_gpu(self, zs, filtered):
        """Post-process + RLE-encode the block's slices in two phases: (A) enqueue every slice's
        kernels on the current stream with no host synchronisation — one emp_stack_slice call per
        slice; run / instance tables and the three status blocks of every slice stay in HBM — then ONE
        synchronisation (the status read-back); (B) read the tables back in two bulk copies and
        assemble the reference's nested dicts on the host.  A slice that overflowed a table (more
        centers or runs than the deferred capacities) is simply redone synchronously."""
        from empanada_b200 import _cabi as C
        from empanada_b200.inference import postprocess as pp
        e = self.engine
        n = len(zs)
        dev = filtered[zs[0]].device
        H, W = filtered[zs[0]].shape[-2:]
        run_cap = max(1 << 14, (H * W * self.upsampling * self.upsampling) // 256)
        inst_cap = max(1 << 12, run_cap // 4)
        runs_all = torch.empty((n, run_cap, 3), dtype=torch.int64, device=dev)
        inst_all = torch.empty((n, inst_cap, 8), dtype=torch.int64, device=dev)
        status_dev = torch.zeros((n, 3, C.ST_WORDS), dtype=torch.int32, device=dev)
        # ---- phase A: one emp_stack_slice call per slice (harden, coarse ids, merge, crop, RLE tables, status gather)
        L = C.lib()
        things, nt = C.i64_array(e.thing_list)
        labels, nl = C.i64_array(self.labels)
        step = 4 if e.coarse_boundaries else 1
        s_up = int(self.upsampling * step)
        shift = int(math.log2(s_up))
        assert (1 << shift) == s_up

        def f32c(t):
            t = t.detach()
            return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()

        # The kernels of one 2048^2 slice are small (~25 launches of a few microseconds each), so slices go round-robin
        # over a few side streams and overlap on the GPU; every stream has its own scratch (C.workspace is per stream).
        main = torch.cuda.current_stream(dev)
        self._streams = _side_streams(dev, self.n_streams)
        ready = torch.cuda.Event()
        ready.record(main)                                  # the median chain's planes and the tables above exist
        for sd in self._streams:
            sd.wait_event(ready)
        t_a = time.perf_counter()
        with torch.cuda.device(dev):
            for i, z in enumerate(zs):
                h = self.heads[z]
                sd = self._streams[i % len(self._streams)]
                with torch.cuda.stream(sd):                 # conversions (if any), scratch and kernels all on the side stream
                    sem, hm, off = f32c(filtered[z]), f32c(h['ctr_hmp']), f32c(h['offsets'])
                    C.require_cuda(sem, hm, off)
                    assert sem.size(0) == 1 and hm.size(0) == 1 and off.size(0) == 1
                    cc, Hs, Ws = sem.shape[1:]
                    hh, ww = hm.shape[-2:]
                    crop = (Hs, Ws) if h['size'] is None else (min(int(h['size'][0]), Hs), min(int(h['size'][1]), Ws))
                    k_cap = min(pp.DEFAULT_K_CAP, hh * ww)
                    nbytes = L.emp_stack_slice_scratch_bytes(Hs, Ws, hh, ww, k_cap, nt, run_cap, nl, int(e.label_divisor))
                    if nbytes == 0:
                        raise ValueError('bad arguments to emp_stack_slice')
                    scratch = C.workspace(dev, nbytes, 'stack_slice')
                    C.check(L.emp_stack_slice(sem.data_ptr(), cc, Hs, Ws, float(e.confidence_thr), hm.data_ptr(), off.data_ptr(),
                                              hh, ww, float(e.nms_threshold), int(e.nms_kernel), float(step), shift, things, nt,
                                              int(e.label_divisor), int(e.stuff_area), int(e.void_label), k_cap, crop[0], crop[1],
                                              labels, nl, int(bool(self.force_connected)), scratch.data_ptr(), scratch.numel(),
                                              None, runs_all[i].data_ptr(), run_cap, inst_all[i].data_ptr(), inst_cap,
                                              status_dev[i].data_ptr(), ctypes.c_void_p(sd.cuda_stream)))
        for sd in self._streams:                            # the block is complete when every side stream is
            done = torch.cuda.Event()
            done.record(sd)
            main.wait_event(done)
        t_b = time.perf_counter()
        status = status_dev.cpu()                           # the one synchronisation of the block
        t_c = time.perf_counter()
        st = status.numpy()                                 # ---- phase B: read back, assemble
        n_runs = st[:, 2, C.ST_NRUNS].astype(np.int64)
        n_inst = st[:, 2, C.ST_NINST].astype(np.int64)
        bad = ((st[:, 0, C.ST_FLAGS] & C.FLAG_K_OVERFLOW) != 0) | ((st[:, 2, C.ST_FLAGS] & C.FLAG_RLE_OVERFLOW) != 0)
        for f in st[:, 1, C.ST_FLAGS]:
            pp._check_flags(int(f))
        ok = ~bad
        mr = int(n_runs[ok].max()) if ok.any() else 0
        mi = int(n_inst[ok].max()) if ok.any() else 0
        # group every slice's runs by instance slot with ONE device sort over the block:
        # key = slice << 44 | slot << 24 | position (runs are in ascending start order already)
        nr_dev = torch.from_numpy(np.where(ok, n_runs, 0)).to(dev)
        rv = runs_all[:, :max(mr, 1)]
        pos = torch.arange(rv.shape[1], device=dev, dtype=torch.int64)
        key = (torch.arange(n, device=dev, dtype=torch.int64)[:, None] << 44) | (rv[:, :, 2] << 24) | pos[None]
        key = torch.where(pos[None] < nr_dev[:, None], key, torch.full_like(key, torch.iinfo(torch.int64).max))
        perm = torch.sort(key.reshape(-1)).indices[:int(nr_dev.sum())]
        starts_h = rv[:, :, 0].reshape(-1)[perm].cpu().numpy()
        lens_h = rv[:, :, 1].reshape(-1)[perm].cpu().numpy()
        inst_h = inst_all[:, :max(mi, 1)].cpu().numpy()
        segs, slot_areas, at = {}, [None] * n, 0
        with _gc_paused():
            self._assemble(zs, segs, slot_areas, bad, n_runs, n_inst, inst_h, starts_h, lens_h, filtered)
        # host seconds: enqueueing, waiting for the device, read-back + dict assembly
        self.timing_ = {'enqueue_s': t_b - t_a, 'wait_s': t_c - t_b, 'assemble_s': time.perf_counter() - t_c}
        self.tables_shape_ = (int(H), int(W))
        # kept for match(): the run tables stay in HBM, slots index each slice's instances in dict order
        self.tables_ = {'runs_all': runs_all, 'n_runs': np.where(ok, n_runs, 0), 'bad': bad,
                        'inst': [inst_h[i, :n_inst[i]] if ok[i] else None for i in range(n)], 'zs': list(zs),
                        'slot_areas': slot_areas}
        return segs

    def _assemble(self, zs, segs, slot_areas, bad, n_runs, n_inst, inst_h, starts_h, lens_h, filtered):
        """Phase B's host loop: the reference's nested dicts per slice from the grouped run tables."""
        from empanada_b200.inference import rle
        e = self.engine
        at = 0
        for i, z in enumerate(zs):
            if bad[i]:
                segs[z] = _slice_sync(e, self.heads[z], filtered[z], self.labels, self.upsampling, self.force_connected)
            else:
                k = int(n_runs[i])
                ins = inst_h[i, :n_inst[i]]
                segs[z] = rle.grouped_to_rle_seg(ins, starts_h[at:at + k], lens_h[at:at + k], self.labels)
                if ins.shape[0]:                        # pixels per instance slot (runs are grouped by slot)
                    first = np.concatenate(([0], np.cumsum(ins[:-1, 6])))
                    slot_areas[i] = np.add.reduceat(lens_h[at:at + k], first) if k else np.zeros(0, np.int64)
                else:
                    slot_areas[i] = np.zeros(0, np.int64)
                at += k

    def _class_overlaps(self, pair_rows, inst_a, inst_b, c):
        """Rows of one slice pair restricted to class c, slots renumbered within the class."""
        sa, sb, ov = pair_rows
        ca, cb = inst_a[:, 0] == c, inst_b[:, 0] == c
        fa = int(np.argmax(ca)) if ca.any() else 0
        fb = int(np.argmax(cb)) if cb.any() else 0
        k = ca[sa] & cb[sb]
        return sa[k] - fa, sb[k] - fb, ov[k]

    def match(self, segs, merge_iou_thr=0.25, merge_ioa_thr=0.25):
        """Forward + backward cross-slice matching (the host loops of patterns.forward_matching /
        backward_matching, patterns.py:68-112, per thing class), with every IoU / IoA taken from ONE
        overlap launch over the block's run tables (inference/matcher.py).  segs: what finish()
        returned.  Returns {z: matched rle_seg}.

        With several ranks the chains are inherently sequential in z: rank r continues the forward chain
        from the state rank r-1 hands over (its last slice's matched instances, label counter and run
        table, a few hundred KB through torch.distributed's object send/recv), and the backward chain
        from rank r+1's; only the boundary pair's overlaps are computed on top of the block's own."""
        with _gc_paused():
            return self._match(segs, merge_iou_thr, merge_ioa_thr)

    def _match(self, segs, merge_iou_thr, merge_ioa_thr):
        from empanada_b200.inference import matcher as mt
        e = self.engine
        t = self.tables_
        zs = t['zs']
        assert zs == sorted(segs.keys())
        assert not t['bad'].any() or self.world == 1, 'multi-rank matching needs the deferred tables of every slice'
        things = [c for c in self.labels if c in e.thing_list]
        dev = t['runs_all'].device
        multi = self.world > 1 and dist.is_available() and dist.is_initialized()

        def exchange(obj, dst, src):
            """send obj to dst (if any), receive from src (if any) — blocking, object collectives"""
            got = None
            if src is not None:
                box = [None]
                dist.recv_object_list(box, src=src, group=self.group, device=dev)
                got = box[0]
            if dst is not None:
                dist.send_object_list([obj], dst=dst, group=self.group, device=dev)
            return got

        if t['bad'].any():                              # a slice was redone synchronously: use the dict API
            pair_rows = None
        else:
            pair_rows = mt.block_overlaps(t['runs_all'], t['n_runs'])
        per_class = {}
        for c in things:
            rles = [segs[z][c] for z in zs]
            if pair_rows is None:
                overlaps, areas = [], None
                for a, b in zip(rles[:-1], rles[1:]):
                    _, _, sa, ra = mt.unpack_rle_attrs(a)
                    _, _, sb, rb = mt.unpack_rle_attrs(b)
                    m = mt.pair_overlaps(sa, ra, sb, rb)
                    i, j = np.nonzero(m)
                    overlaps.append((i, j, m[i, j]))
            else:
                areas = [t['slot_areas'][i][t['inst'][i][:, 0] == c] for i in range(len(zs))]
                overlaps = [self._class_overlaps(rows, t['inst'][p], t['inst'][p + 1], c) for p, rows in enumerate(pair_rows)]
            per_class[c] = (rles, overlaps, areas, mt.StackMatcher(c, e.label_divisor, merge_iou_thr, merge_ioa_thr))

        # ---- forward: continue from the rank below, hand over to the rank above
        prev = None
        if multi and self.rank > 0:
            prev = exchange(None, None, self.rank - 1)
            # overlaps of the boundary pair (their last slice, my first slice) from the two run tables
            mine = t['runs_all'][0, :int(t['n_runs'][0])]
            theirs = torch.from_numpy(prev['runs']).to(dev)
            stride = max(int(mine.shape[0]), int(theirs.shape[0]), 1)
            both = torch.zeros((2, stride, 3), dtype=torch.int64, device=dev)
            both[0, :theirs.shape[0]] = theirs
            both[1, :mine.shape[0]] = mine
            rows = mt.block_overlaps(both, [int(theirs.shape[0]), int(mine.shape[0])])[0]
            self._boundary_below = {c: self._class_overlaps(rows, prev['inst'], t['inst'][0], c) for c in things}
        fwd = {}
        for c, (rles, overlaps, areas, sm) in per_class.items():
            st = None
            if prev is not None:
                st = dict(prev['classes'][c], overlaps=self._boundary_below[c])
            fwd[c] = sm.forward(rles, overlaps, areas, prev=st)
        if multi and self.rank + 1 < self.world:
            last = len(zs) - 1
            exchange({'runs': t['runs_all'][last, :int(t['n_runs'][last])].cpu().numpy(), 'inst': t['inst'][last],
                      'classes': {c: per_class[c][3].forward_state(fwd[c][1]) for c in things}}, self.rank + 1, None)

        # ---- backward: continue from the rank above, hand over to the rank below
        nxt = exchange(None, None, self.rank + 1) if multi and self.rank + 1 < self.world else None
        out = {z: dict(segs[z]) for z in zs}
        # final label of every instance slot (all classes): its own label unless a matcher renames it
        self.slot_labels_ = [None if t['inst'][i] is None else t['inst'][i][:, 1].copy() for i in range(len(zs))]
        down = {}
        for c, (rles, overlaps, areas, sm) in per_class.items():
            bwd = sm.backward(fwd[c][0], fwd[c][1], rles, overlaps, nxt=None if nxt is None else nxt[c])
            down[c] = sm.backward_state(bwd)
            for i, (z, seg) in enumerate(zip(zs, bwd)):
                out[z][c] = seg
                if self.slot_labels_[i] is not None:
                    self.slot_labels_[i][t['inst'][i][:, 0] == c] = sm.slot_labels[i]
        if multi and self.rank > 0:
            # the rank below matches its last slice against my first: overlaps as (its slot, my slot, inter)
            exchange({c: dict(down[c], overlaps=self._boundary_below[c]) for c in things}, self.rank - 1, None)
        return out

    def fill(self, dtype=torch.int64):
        """The block as labelled planes (n, H, W) in HBM, painted from the run tables with the labels the
        last match() assigned (finish() labels if match() was not called): the GPU counterpart of
        array_utils.numpy_fill_instances over the tracker's instances (array_utils.py:725-736)."""
        from empanada_b200.inference import fill as fl
        t = self.tables_
        assert not t['bad'].any(), 'a slice was redone synchronously: fill from the dicts instead (fill_instances)'
        labels = getattr(self, 'slot_labels_', None) or [ins[:, 1] for ins in t['inst']]
        width = max([len(l) for l in labels] + [1])
        table = np.full((len(labels), width), -1, np.int64)
        for i, l in enumerate(labels):
            table[i, :len(l)] = l
        h = self.heads[t['zs'][0]]
        H, W = h['size'] if h['size'] is not None else self.tables_shape_
        return fl.fill_block(t['runs_all'], t['n_runs'], table, (H, W), dtype)

    def _all_ranks_agree(self, flag, device):
        """True iff `flag` holds on every rank (the fast carry hand-over needs all ranks to take it)."""
        t = torch.tensor([int(bool(flag))], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(t.item())

    def finish(self):
        from empanada_b200.inference import engines as eng
        from empanada_b200.inference import rle

        raw = {z: h['sem'] for z, h in self.heads.items()}

        def med(window):
            return eng.median_harden(window, 0.0)[0]

        def chain(carry):
            return median_chain(raw, self.z0, self.z1, self.depth, self.ks, carry, med)

        def compose():
            """(A, B) of this rank's block from emp_median3_compose over its raw planes (+ the first halo plane)."""
            from empanada_b200 import _cabi as C
            last_raw = self.z1 >= self.depth
            planes = [raw[z].detach() for z in range(self.z0, self.z1 if last_raw else self.z1 + 1)]
            assert all(p.dtype == torch.float32 and p.is_contiguous() and p.shape == planes[0].shape for p in planes)
            dev = C.require_cuda(*planes)
            ptrs = torch.tensor([p.data_ptr() for p in planes], dtype=torch.int64).to(dev)
            A, B = torch.empty_like(planes[0]), torch.empty_like(planes[0])
            with torch.cuda.device(dev):
                C.check(C.lib().emp_median3_compose(ptrs.data_ptr(), self.z1 - self.z0, int(self.z0 == 0), int(last_raw),
                                                    planes[0].numel(), A.data_ptr(), B.data_ptr(), C.stream_ptr(dev)))
            return A, B

        # ks == 3 over several ranks: the carry crosses the ranks as one clamp per rank (no rank waits for the chain of
        # the rank below); anything else — and the NaN fallback — hands the carry on after each rank's chain
        mismatch = None
        fast = (self.ks == 3 and self.world > 1 and self.z1 > self.z0 and not getattr(self, '_sequential_carry', False)
                and all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in raw.values()))
        fast = self._all_ranks_agree(fast, raw[self.z0].device) if self.ks == 3 and self.world > 1 else False
        if fast:
            filtered, mismatch = exchange_carry_median3(compose, chain, self.rank, self.world, raw[self.z0], self.group)
        else:
            filtered = exchange_carry(chain, self.rank, self.world, self.mid, raw[self.z0], self.group)
        e = self.engine
        zs = list(range(self.z0, self.z1))
        if not zs:
            segs = {}
        elif filtered[self.z0].is_cuda:
            with _gc_paused():
                segs = self._finish_block_gpu(zs, filtered)
        else:
            raise RuntimeError('StackShard runs on CUDA tensors only (there is no CPU fallback)')
        out, max_counts = {}, {c: 0 for c in self.labels}
        for z in zs:
            seg = segs[z]
            for c in self.labels:
                if c in e.thing_list and seg[c]:
                    max_counts[c] = max(max_counts[c], max(seg[c]) - c * e.label_divisor)
            out[z] = seg
        counts = torch.tensor([max_counts[c] for c in self.labels] + [0], dtype=torch.int64, device=raw[self.z0].device)
        if mismatch is not None:
            counts[-1] = mismatch                           # rides along with the instance counts
        # a single-rank shard never talks to anyone, even inside a larger process group
        if self.world == 1:
            offs = torch.zeros_like(counts[:-1])
        else:
            offs, table = label_offsets(counts, self.group)
            if bool(table[:, -1].any()):                    # some rank's composed carry was not its chain's plane (NaNs):
                self._sequential_carry = True               # every rank sees the same table, so all of them redo the block
                return self.finish()
            offs = offs[:-1]
        offs = {c: int(o) for c, o in zip(self.labels, offs.tolist())}
        self.label_offsets_ = offs
        return {z: apply_label_offset(s, offs, e.label_divisor, e.thing_list) for z, s in out.items()}
