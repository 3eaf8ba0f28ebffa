"""
Z-sharded stack inference: the part of the reference's 3D path between "CNN heads produced
sem / ctr_hmp / offsets for slice z" and "the host tracker receives slice z's RLE dict"
(reference scripts/pdl_inference3d.py:163-176 + empanada/inference/patterns.py:68-100, and the
multi-GPU recipe patterns.forward_multigpu :279-350), one process per GPU.

Sharding.  Slices are independent except for the recursive median queue (engines.py:47-90):
    m_i = median(f_{i-mid}, ..., f_{i-1}, s_i, s_{i+1}, ..., s_{i+mid}),   f_j = m_j (j >= mid) else s_j
for mid <= i < D - mid, raw s_i otherwise.  Rank r owns the contiguous block [z0, z1).  It needs
  * a look-ahead halo: the raw probabilities of slices z1 .. z1+mid-1 (it runs the CNN on them too);
  * a carry: the filtered planes f_{z0-mid} .. f_{z0-1} of rank r-1.
No rank waits for the chain of the rank below.  Every rank runs its whole chain at once from a GUESSED carry
(its own first raw plane; one emp_median_chain launch, each probability read once), sends the `mid` planes it
ends with to rank r+1 (one NCCL send/recv per neighbour pair over NVLink, all pairs at the same time) and then
REPAIRS its block from the carry it received: per pixel the guessed and the true chain advance together until
their states agree bit for bit — an order-statistic filter forgets its start within a few slices — and only that
prefix is redone (emp_median_chain_repair).  A rank whose repair reached the end of its block says so; the flag
rides along with the instance counts, and only then (never seen on EM data; NaNs force it) the ranks settle with
further rounds, each one correcting at least one more rank, before the block is post-processed again.

Blocks.  Behind the chain a z-block is processed `block` slices at a time: ONE emp_stack_block call — a dozen
launches — per sub-block, its run tables packed into one buffer that crosses PCIe in one copy while the next
sub-block runs.  The int64 label map is never materialised; the dicts of the reference's format are built lazily
from the packed tables (RleStack).

Labels.  Every slice numbers its instances 1..n per class.  So that labels from different ranks
never collide before the host-side cross-slice matcher renumbers them, ranks all-gather their
per-class maximum instance count (one int64 per class) and add the exclusive prefix as an offset;
rank 0 keeps offset 0, which is what the reference's matcher sees for the first slice.
"""
import collections.abc
import contextlib
import ctypes
import gc
import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

__all__ = ['partition_slices', 'halo_range', 'median_chain', 'exchange_planes', 'carry_rounds', 'label_offsets',
           'apply_label_offset', 'RleStack', 'StackShard']


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


_copy_streams = {}
_pinned = {}
_words_per_slice = {}       # (device, H, W) -> packed words per slice seen so far (sizes the one D2H copy per sub-block)


def _copy_stream(device):
    s = _copy_streams.get(device.index)
    if s is None:
        s = _copy_streams[device.index] = torch.cuda.Stream(device)
    return s


class _PinnedPool:
    """Pinned int64 buffers the packed tables land in.  A block's RleStack reads its arrays IN PLACE (views, no copy:
    copying ~120 KB per slice out of the staging area cost more host time than the GPU needed for the slice), so a
    buffer stays taken for as long as any of those views is alive and goes back to the pool when the last one dies —
    the pool simply looks at reference counts.  Pinning is slow (milliseconds): buffers are kept and over-allocated."""

    def __init__(self):
        self.buffers = {}

    def acquire(self, device, n_words):
        import sys
        have = self.buffers.setdefault(device.index, [])
        for i in range(len(have)):
            t = have[i]
            # references when nobody else holds it: the list slot, `t`, and getrefcount's argument
            if t.numel() >= n_words and sys.getrefcount(t) <= 3:
                return t
        have[:] = [t for t in have if sys.getrefcount(t) > 3 or t.numel() >= n_words][-4:]       # drop idle buffers that are too small
        t = torch.empty(int(n_words * 1.5) + 1024, dtype=torch.int64).pin_memory()
        have.append(t)
        return t


_table_pool = _PinnedPool()


class _DevicePool:
    """Large per-block device tensors (class maps, packed tables, row-run tables) kept from block to block: the caching
    allocator was seen to take ~0.7 ms for a 25 MB torch.empty next to NCCL traffic, a quarter of a 64-slice block.  Like
    the pinned pool, a tensor is free again when nobody but the pool references it."""

    def __init__(self):
        self.tensors = {}

    def acquire(self, device, shape, dtype):
        import sys
        n = 1
        for v in shape:
            n *= int(v)
        have = self.tensors.setdefault((device.index, dtype), [])
        for t in have:
            if t.numel() >= n and t.numel() <= 2 * n + 4096 and sys.getrefcount(t) <= 3:
                return t[:n].view(shape)
        if len(have) > 12:
            have[:] = [t for t in have if sys.getrefcount(t) > 3][-12:]
        t = torch.empty(n, dtype=dtype, device=device)
        have.append(t)
        return t[:n].view(shape)


_device_pool = _DevicePool()


def _pinned_words(device, slot, n_words):
    """Pinned int64 staging buffer `slot` of this device, grown and never shrunk."""
    key = (device.index, slot)
    t = _pinned.get(key)
    if t is None or t.numel() < n_words:
        # twice what is asked for: pinning is slow (milliseconds), and the size guess moves with the data
        t = _pinned[key] = torch.empty(int(2 * n_words), dtype=torch.int64).pin_memory()
    return t


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
    """Recursive median over the block [z0, z1), plane by plane — the restatement of the reference's queue the
    kernels are tested against (tests/test_stack_host.py, tests/test_gpu_stack_block.py).

    raw      dict {z: sem plane} for z in [z0, min(depth, z1 + mid))
    carry    list of the `mid` planes f_{z0-mid} .. f_{z0-1} (filtered where z >= mid, else raw);
             ignored for z0 == 0
    median_fn(list of ks planes) -> plane
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


def exchange_planes(send, recv, rank, world_size, group=None):
    """Every rank hands `send` (list of tensors) to rank + 1 and fills `recv` from rank - 1, all neighbour pairs at
    the same time (one batched NCCL send/recv; gloo on CPU tensors)."""
    ops = []
    if rank + 1 < world_size:
        ops += [dist.P2POp(dist.isend, t, rank + 1, group) for t in send]
    if rank > 0:
        ops += [dist.P2POp(dist.irecv, t, rank - 1, group) for t in recv]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def carry_rounds(rank, world_size, chain, repair, exchange, any_changed=None):
    """The carry hand-over of a z-sharded recursive median without a rank-after-rank wait.

    chain()            run this rank's chain from the guessed carry (rank 0: exact), leaving its outgoing carry
    exchange()         outgoing carry -> rank + 1, rank - 1's -> this rank's `received` buffer
    repair(round)      redo the block's prefix from the received carry against the one used before (round 0: the
                       guess); returns a flag (tensor) that is non-zero when the outgoing carry changed
    any_changed(flag)  None: ONE round, the flag is returned for the caller to share later (the common path: if no
                       rank's flag is set, every block is exact — rank r's repair saw rank r-1's final carry);
                       else a callable reducing the flag over all ranks: rounds repeat until no carry moves (each
                       round settles at least one more rank, so world_size - 1 rounds always suffice)
    """
    chain()
    flag = None
    for rnd in range(max(world_size - 1, 1)):
        exchange()
        flag = repair(rnd) if rank > 0 else None
        if any_changed is None:
            return flag
        if not any_changed(flag):
            return None
    return None


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


class _BlockTables:
    """Host copy of one emp_stack_block packed output (include/empanada_b200.h)."""
    __slots__ = ('slices', 'starts', 'lens', 'inst')

    def __init__(self, words, B):
        from empanada_b200 import _cabi as C
        R, I = int(words[1]), int(words[2])
        Rp = (R + 1) & ~1
        s0 = C.BLK_HDR_WORDS + C.BLK_SLICE_WORDS * B
        # views: `words` is a slice of a pooled pinned buffer that stays taken while these arrays live (_PinnedPool)
        self.slices = words[C.BLK_HDR_WORDS:s0].reshape(B, C.BLK_SLICE_WORDS)
        runs32 = words[s0:s0 + Rp].view(np.int32)              # the run lists cross PCIe as int32 (flat indices < 2^31)
        self.starts = runs32[:R]
        self.lens = runs32[Rp:Rp + R]
        self.inst = words[s0 + Rp:s0 + Rp + C.BLK_INST_WORDS * I].reshape(I, C.BLK_INST_WORDS)

    @staticmethod
    def words_needed(words, B):
        from empanada_b200 import _cabi as C
        return C.BLK_HDR_WORDS + C.BLK_SLICE_WORDS * B + ((int(words[1]) + 1) & ~1) + C.BLK_INST_WORDS * int(words[2])


class RleStack(collections.abc.Mapping):
    """{z: rle_seg} of a z-block in the reference's format ({class: {label: {'box', 'starts', 'runs'}}},
    rle.py:26-86), built per slice on first access from the packed run / instance tables the GPU returned: the
    tables hold everything (instance rows in dict order, each instance's starts / runs contiguous), so consumers
    that work on arrays — the cross-slice matcher, counters, fill — never pay for ~200 small dicts per slice."""

    def __init__(self, labels, thing_list, label_divisor):
        self.labels, self.thing_list, self.label_divisor = [int(l) for l in labels], list(thing_list), int(label_divisor)
        self._where = {}        # z -> (block tables, index in block) | dict (a slice redone synchronously)
        self._cache = {}
        self.offsets = {}

    def _add(self, z, tables, b):
        self._where[z] = (tables, b)

    def _add_dict(self, z, seg):
        self._where[z] = seg

    def __len__(self):
        return len(self._where)

    def __iter__(self):
        return iter(sorted(self._where))

    def inst_rows(self, z):
        """(n, 9) int64 rows of slice z: class, label (before the rank offset), y0, x0, y1, x1, n runs, first run, area;
        None for a slice that was redone synchronously."""
        w = self._where[z]
        if isinstance(w, dict):
            return None
        t, b = w
        n, first = int(t.slices[b, 0]), int(t.slices[b, 1])
        return t.inst[first:first + n]

    def counts(self):
        """(instances, runs) over the block without building a dict."""
        n_inst = n_runs = 0
        for z, w in self._where.items():
            if isinstance(w, dict):
                n_inst += sum(len(v) for v in w.values())
                n_runs += sum(len(a['starts']) for v in w.values() for a in v.values())
            else:
                rows = self.inst_rows(z)
                n_inst += rows.shape[0]
                n_runs += int(rows[:, 6].sum())
        return n_inst, n_runs

    def __getitem__(self, z):
        seg = self._cache.get(z)
        if seg is not None:
            return seg
        w = self._where[z]
        if isinstance(w, dict):
            seg = apply_label_offset(w, self.offsets, self.label_divisor, self.thing_list)
        else:
            t, _ = w
            rows = self.inst_rows(z)
            seg = {l: {} for l in self.labels}
            if rows.shape[0]:
                cls = rows[:, 0].tolist()
                labs = rows[:, 1].tolist()
                boxes = rows[:, 2:6].tolist()
                cnt = rows[:, 6].tolist()
                first = rows[:, 7].tolist()
                # the slice's run lists widened to the reference's int64 once; the instances' arrays are views into them
                lo = first[0]
                hi = first[-1] + cnt[-1]
                starts, lens = t.starts[lo:hi].astype(np.int64), t.lens[lo:hi].astype(np.int64)
                first = [f - lo for f in first]
                offs = {c: (int(self.offsets.get(c, 0)) if c in self.thing_list else 0) for c in self.labels}
                for i in range(len(cls)):
                    c = cls[i]
                    lab = labs[i] + offs[c]
                    if lab >= (c + 1) * self.label_divisor:
                        raise ValueError(f'label {labs[i]} + offset {offs[c]} leaves class {c}\'s range; '
                                         f'raise label_divisor (reference default for 3D is 20000)')
                    a, e = first[i], first[i] + cnt[i]
                    seg[c][lab] = {'box': tuple(boxes[i]), 'starts': starts[a:e], 'runs': lens[a:e]}
        self._cache[z] = seg
        return seg

    def materialise(self):
        """Build every slice's dict now (what a consumer iterating over .items() pays)."""
        with _gc_paused():
            for z in self._where:
                self[z]
        return self


def _slice_sync(engine, head, sem8, labels, upsampling, force_connected):
    """One slice, synchronously (status read-backs and retries inside): the fallback for a slice
    whose deferred run overflowed a table."""
    from empanada_b200.inference import rle
    pan = engine._fused_postprocess(None, head['ctr_hmp'], head['offsets'], upsampling, sem8=sem8)
    if head['size'] is not None:
        pan = pan[..., :head['size'][0], :head['size'][1]]
    return rle.pan_seg_to_rle_seg(pan, labels, engine.label_divisor, engine.thing_list, force_connected)


def _f32c(t):
    # only data pointers are taken from these tensors (no autograd graph is touched): the common case costs two checks
    if t.dtype is torch.float32 and t.is_contiguous():
        return t
    return t.detach().to(torch.float32).contiguous()


def _same_planes(planes):
    """All planes on one CUDA device, (1, C, H, W), equal in shape — checked cheaply: a block has hundreds of them and
    this runs before the first launch."""
    from empanada_b200 import _cabi as C
    p0 = planes[0]
    dev = C.require_cuda(p0, planes[-1])
    assert p0.dim() == 4 and p0.size(0) == 1 and planes[-1].shape == p0.shape
    idx, numel = p0.get_device(), p0.numel()
    for p in planes:
        if p.get_device() != idx or p.numel() != numel:
            raise RuntimeError('the slices of a block must be CUDA tensors of one device and one shape')
    return dev


def _batched(tensors):
    """(base tensor, stride in elements) of equal-shape contiguous tensors: zero-copy when they already sit evenly
    spaced in one allocation (views of a batch), else one torch.stack."""
    t0 = tensors[0]
    n = len(tensors)
    if n == 1:
        return t0, t0.numel()
    step = tensors[1].data_ptr() - t0.data_ptr()
    if step >= t0.numel() * 4 and step % 16 == 0 and all(tensors[i].data_ptr() - t0.data_ptr() == i * step for i in range(2, n)) \
            and t0.untyped_storage().data_ptr() == tensors[-1].untyped_storage().data_ptr():
        return t0, step // 4
    st = torch.stack([t.reshape(-1) for t in tensors])
    return st, st.shape[1]


class StackShard:
    """One rank's share of a stack: feed it the head tensors of its slices (block + halo) in z
    order, then `finish()` returns {z: rle_seg} (an RleStack) for the block.

    engine: a PanopticDeepLabRenderEngine(3d) (its post-processing parameters and fused kernels
    are used; its own median queue is bypassed in favour of the sharded chain above).
    block: slices per emp_stack_block call (sub-block); keep_tables: keep the (start, length, slot) row-run tables
    in HBM for match() / fill().
    stream: an optional torch.cuda.Stream for advance(): the streamed sub-blocks then run beside the producer of the
    heads (the CNN on the current stream) instead of between its kernels.
    """

    def __init__(self, engine, labels, depth, rank=0, world_size=1, median_kernel_size=3,
                 upsampling=1, force_connected=True, group=None, block=128, keep_tables=True, chain_chunk=4096, run_cap=None, inst_cap=None,
                 stream=None):
        assert median_kernel_size % 2 == 1, "Kernel size must be odd integer!"
        assert math.log(upsampling, 2).is_integer(), "Upsampling factor not log base 2!"
        self.engine, self.labels, self.depth = engine, list(labels), depth
        self.rank, self.world, self.ks = rank, world_size, median_kernel_size
        self.mid = (median_kernel_size - 1) // 2
        if depth < median_kernel_size:
            raise ValueError(f'stack of {depth} slices is shallower than the median kernel ({median_kernel_size})')
        if world_size > 1 and depth // world_size < max(self.mid, 1):
            # every rank sees the same arguments, so every rank raises: nobody is left waiting in a collective
            raise ValueError(f'{depth} slices over {world_size} ranks leaves blocks shorter than {max(self.mid, 1)} slice(s); '
                             f'use fewer ranks')
        self.upsampling, self.force_connected, self.group = upsampling, force_connected, group
        self.z0, self.z1 = partition_slices(depth, world_size, rank)
        _, self.z_halo = halo_range(depth, world_size, rank, median_kernel_size)
        self.heads = {}
        self.block = max(1, int(block))
        self.keep_tables = bool(keep_tables)
        self.chain_chunk = max(1, int(chain_chunk))      # slices per emp_median_chain launch
        self.run_cap, self.inst_cap = run_cap, inst_cap  # deferred table capacities per slice (None: from the plane size)
        self._settle = False
        self._stream = None                               # None: not begun; False: not streamable; dict: state of advance()
        self.side_stream = stream                         # advance() enqueues here (None: on the current stream)
        self._geom = None

    def slices(self):
        """z indices this rank must run the CNN on, in order."""
        return range(self.z0, self.z_halo)

    def add(self, z, sem_prob, ctr_hmp=None, offsets=None, size=None):
        """Head tensors of slice z: sem_prob (1,C,H,W) probabilities; ctr_hmp/offsets only needed
        for z inside the block (halo slices contribute their probabilities only)."""
        assert self.z0 <= z < self.z_halo
        self.heads[z] = {'sem': sem_prob, 'ctr_hmp': ctr_hmp, 'offsets': offsets, 'size': size}

    # ------------------------------------------------------------------------------------------------------
    def _geometry(self, H, W):
        """(coarse h, w, shift) of the block: the coarse maps are 2^shift times smaller than the H x W planes."""
        if self._geom is not None:                                  # streamed blocks forget their first slices' heads
            return self._geom
        hh, ww = self.heads[self.z0]['ctr_hmp'].shape[-2:]
        s_up = int(self.upsampling * (4 if self.engine.coarse_boundaries else 1))
        shift = int(math.log2(s_up))
        assert (1 << shift) == s_up
        self._geom = (int(hh), int(ww), shift)
        return self._geom

    def _chain(self, planes, dev, hw, Cn, need=None):
        """Recursive median + harden over the rank's block: sem8 (n, hw) uint8 in HBM.  Multi-rank blocks start from a
        guessed carry and are repaired from the true one (module docstring).  need: optional (n, h*w) uint8 zeros the
        kernels mark with the coarse cells that hold a thing pixel.  Returns (sem8, changed flag or None)."""
        from empanada_b200 import _cabi as C
        e, L = self.engine, C.lib()
        n, mid, ks = self.z1 - self.z0, self.mid, self.ks
        stream = C.stream_ptr(dev)
        vp = ctypes.c_void_p
        sem8 = _device_pool.acquire(dev, (n, hw), torch.uint8)
        need_arg = need_map = None
        if need is not None:
            _, ww, shift = self._geometry(*self._plane)
            bits = 0
            for c in e.thing_list:
                bits |= 1 << int(c)
            need_map = C.NeedMap(map=need.data_ptr(), stride=need.shape[1], W=self._plane[1], shift=shift, wc=ww, thing_bits=bits)
            need_arg = ctypes.byref(need_map)
        multi = self.world > 1 and mid > 0 and dist.is_available() and dist.is_initialized()
        words = [p.data_ptr() for p in planes]
        carry = {}
        names = ('out', 'a', 'b', 'p0', 'p1') if multi else ('out', 'p0', 'p1')
        if mid > 0:
            for name in names:
                carry[name] = [torch.empty((Cn * hw,), dtype=torch.float32, device=dev) for _ in range(mid)]
                words += [t.data_ptr() for t in carry[name]]
            words += [planes[0].data_ptr()] * mid                      # the guess: this rank's first raw plane
        tab = torch.tensor(words, dtype=torch.int64).to(dev)
        base = tab.data_ptr()
        at = {'planes': base}
        o = len(planes)
        for name in names + ('guess',):
            at[name] = base + 8 * o
            o += mid
        best = torch.empty((n, hw), dtype=torch.float32, device=dev) if Cn > 1 else None
        thr = float(e.confidence_thr)
        self._keep = (tab, carry, best)                                 # alive until the stream has run
        chunk = max(self.chain_chunk, mid, 1)

        def chain(carry_in):
            """The block's chain, `chain_chunk` slices per launch (the filter state crosses launches through `mid` planes).
            One launch over the whole block is fastest when every head is already in HBM (measured on 512 slices:
            3.30 ms whole, 3.56 ms in chunks of 32, 4.26 ms in chunks of 8); chunks are for heads that arrive while the
            CNN is still running."""
            cur = carry_in
            with torch.cuda.device(dev):
                for ci, i0 in enumerate(range(0, n, chunk)):
                    m = min(chunk, n - i0)
                    nxt = None
                    if mid > 0:
                        nxt = at['out'] if i0 + m >= n else at['p0' if ci % 2 == 0 else 'p1']
                    nm = None
                    if need_map is not None:
                        nm = C.NeedMap(map=need_map.map + i0 * need_map.stride, stride=need_map.stride, W=need_map.W,
                                       shift=need_map.shift, wc=need_map.wc, thing_bits=need_map.thing_bits)
                    C.check(L.emp_median_chain(vp(at['planes'] + 8 * i0), m, min(len(planes) - i0, m + mid), self.z0 + i0, self.depth,
                                               ks, Cn, hw, vp(cur) if cur else None, thr, vp(sem8.data_ptr() + i0 * hw), hw,
                                               vp(best.data_ptr() + 4 * i0 * hw) if best is not None else None,
                                               vp(nxt) if nxt else None, ctypes.byref(nm) if nm is not None else None, stream))
                    cur = nxt

        if not multi:
            chain(None)
            return sem8, None
        if Cn > 1:
            # multi-channel blocks keep a running arg-max the repair cannot redo: hand the carry on rank after rank
            if self.rank > 0:
                for t in carry['a']:
                    dist.recv(t, src=self.rank - 1, group=self.group)
            chain(at['a'] if self.rank > 0 else None)
            if self.rank + 1 < self.world:
                for t in carry['out']:
                    dist.send(t, dst=self.rank + 1, group=self.group)
            return sem8, None

        changed = torch.zeros((1,), dtype=torch.int32, device=dev)
        state = {'old': 'guess', 'new': 'a'}

        def repair(rnd):
            changed.zero_()
            with torch.cuda.device(dev):
                C.check(L.emp_median_chain_repair(vp(at['planes']), n, len(planes), self.z0, self.depth, ks, hw,
                                                  vp(at[state['old']]), vp(at[state['new']]), thr, vp(sem8.data_ptr()), hw,
                                                  vp(at['out']), vp(changed.data_ptr()), need_arg, stream))
            state['old'], state['new'] = state['new'], ('b' if state['new'] == 'a' else 'a')
            return changed

        def exchange():
            t = time.perf_counter()
            exchange_planes(carry['out'], carry[state['new']], self.rank, self.world, self.group)
            self._marks['exchange_call_s'] = self._marks.get('exchange_call_s', 0.0) + time.perf_counter() - t

        def any_changed(flag):
            t = (flag if flag is not None else torch.zeros((1,), dtype=torch.int32, device=dev)).to(torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            return bool(t.item())

        if os.environ.get('EMP_STACK_DEBUG_SYNC'):                     # diagnostic: device time of each step of the hand-over
            def timed(name, fn):
                def run(*a):
                    torch.cuda.synchronize(dev)
                    t = time.perf_counter()
                    r = fn(*a)
                    torch.cuda.synchronize(dev)
                    self._marks['dbg_' + name] = self._marks.get('dbg_' + name, 0.0) + time.perf_counter() - t
                    return r
                return run
            chain, repair, exchange = timed('chain_s', chain), timed('repair_s', repair), timed('exchange_s', exchange)
        flag = carry_rounds(self.rank, self.world, lambda: chain(at['guess'] if self.rank > 0 else None), repair, exchange,
                            any_changed if self._settle else None)
        return sem8, flag

    def _blocks_prepare(self, zs, dev, H, W):
        """Everything the block launches of this rank's z-block share: the configuration, the device scratch and packed
        tables, the pinned landing area with one arrival flag per sub-block."""
        from empanada_b200 import _cabi as C
        from empanada_b200.inference import postprocess as pp
        e, L = self.engine, C.lib()
        n = len(zs)
        h0 = self.heads[zs[0]]
        hh, ww, shift = self._geometry(H, W)
        size = h0['size']
        self._size0 = size
        crop = (H, W) if size is None else (min(int(size[0]), H), min(int(size[1]), W))
        step = 4 if e.coarse_boundaries else 1
        things, nt = C.i64_array(e.thing_list)
        labels, nl = C.i64_array(self.labels)
        # a slice that overflows either table is redone synchronously with tables grown to fit (_slice_sync)
        run_cap = int(self.run_cap) if self.run_cap else max(1 << 14, (H * W) // 256)
        inst_cap = int(self.inst_cap) if self.inst_cap else max(1 << 12, run_cap // 4)
        cfg = C.StackCfg(H=H, W=W, h=hh, w=ww, shift=shift, nms_kernel=int(e.nms_kernel), k_cap=min(pp.DEFAULT_K_CAP, hh * ww),
                         crop_h=crop[0], crop_w=crop[1], n_things=nt, n_labels=nl, force_connected=int(bool(self.force_connected)),
                         run_cap=run_cap, inst_cap=inst_cap, nms_threshold=float(e.nms_threshold), step=float(step),
                         label_divisor=int(e.label_divisor), stuff_area=int(e.stuff_area), void_label=int(e.void_label),
                         thing_list=ctypes.cast(things, ctypes.c_void_p), labels=ctypes.cast(labels, ctypes.c_void_p))
        SB = min(self.block, n)
        n_sub = (n + SB - 1) // SB
        packed_words = int(L.emp_stack_block_packed_words(ctypes.byref(cfg), SB))
        scratch_bytes = int(L.emp_stack_block_scratch_bytes(ctypes.byref(cfg), SB))
        if packed_words == 0 or scratch_bytes == 0:
            raise ValueError('bad arguments to emp_stack_block: ' + L.emp_last_error().decode(errors='replace'))
        scratch = C.workspace(dev, scratch_bytes, 'stack_block')
        packed = _device_pool.acquire(dev, (n_sub, packed_words), torch.int64)
        runs_all = _device_pool.acquire(dev, (n, run_cap, 3), torch.int64) if self.keep_tables else None
        maxlab = torch.zeros((C.MAX_LABELS,), dtype=torch.int64, device=dev)
        # one pinned buffer: n_sub table areas of host_words each (sized from what the data needed so far), then n_sub flags
        fixed = C.BLK_HDR_WORDS + C.BLK_SLICE_WORDS * SB
        per_slice = _words_per_slice.get((dev.index, H, W), 1 << 13)
        host_words = min(packed_words, fixed + SB * int(per_slice * 1.25))
        host = _table_pool.acquire(dev, n_sub * (host_words + 1))
        flags = host[n_sub * host_words:n_sub * host_words + n_sub]
        flags.zero_()
        return {'cfg': cfg, 'n': n, 'n_sub': n_sub, 'SB': SB, 'H': H, 'W': W, 'packed_words': packed_words, 'scratch': scratch,
                'packed': packed, 'runs_all': runs_all, 'run_cap': run_cap, 'maxlab': maxlab, 'nl': nl, 'host': host,
                'host_words': host_words, 'flags_t': flags, 'flags': flags.numpy(), 'keep': [things, labels]}

    def _blocks_launch(self, prep, zs, sem8, need, i0=0, m=None):
        """ONE emp_stack_blocks call for slices [i0, i0 + m) of the z-block (default: all of it): per sub-block a dozen
        launches on the current stream and one device->host copy of its packed tables on the copy stream.  No host
        synchronisation.  i0 must be a multiple of the sub-block size."""
        from empanada_b200 import _cabi as C
        L = C.lib()
        n = prep['n']
        m = n - i0 if m is None else m
        SB, H, W = prep['SB'], prep['H'], prep['W']
        assert i0 % SB == 0 and 0 < m <= n - i0
        bi0 = i0 // SB
        dev = sem8.device
        zs_g = zs[i0:i0 + m]
        assert all(self.heads[z]['size'] == self._size0 for z in zs_g), 'slices of one block must share their unpadded size'
        hm, hm_stride = _batched([_f32c(self.heads[z]['ctr_hmp']) for z in zs_g])        # one gather for the whole block
        off, off_stride = _batched([_f32c(self.heads[z]['offsets']) for z in zs_g])
        C.require_cuda(hm, off)
        main = torch.cuda.current_stream(dev)
        side = _copy_stream(dev)
        vp = ctypes.c_void_p
        packed, runs_all, host, hw_ = prep['packed'], prep['runs_all'], prep['host'], prep['host_words']
        scratch = prep['scratch']
        with torch.cuda.device(dev):
            C.check(L.emp_stack_blocks(ctypes.byref(prep['cfg']), m, SB, vp(sem8.data_ptr() + i0 * H * W), H * W, vp(hm.data_ptr()), hm_stride,
                                       vp(off.data_ptr()), off_stride,
                                       vp(need.data_ptr() + i0 * need.shape[1]) if need is not None else None,
                                       need.shape[1] if need is not None else 0, vp(scratch.data_ptr()), scratch.numel(),
                                       vp(packed.data_ptr() + 8 * bi0 * prep['packed_words']), prep['packed_words'],
                                       vp(runs_all.data_ptr() + 8 * 3 * prep['run_cap'] * i0) if runs_all is not None else None,
                                       vp(prep['maxlab'].data_ptr()), vp(host.data_ptr() + 8 * bi0 * hw_), hw_, hw_,
                                       vp(prep['flags_t'].data_ptr() + 8 * bi0), vp(main.cuda_stream), vp(side.cuda_stream)))
        prep['keep'].append((hm, off))

    def _enqueue_blocks(self, zs, sem8, dev, H, W, need=None):
        """Phase A: the whole z-block in one emp_stack_blocks call."""
        t_0 = time.perf_counter()
        prep = self._blocks_prepare(zs, dev, H, W)
        t_1 = time.perf_counter()
        self._blocks_launch(prep, zs, sem8, need)
        self._marks.update(blocks_alloc_s=t_1 - t_0, blocks_pinned_s=0.0, blocks_c_call_s=time.perf_counter() - t_1)
        return prep

    def _collect(self, zs, subs, packed, sem8, out, streamed=False):
        """Phase B: as each block's tables land in pinned memory (its flag word turns non-zero) take owned copies of them and
        register the slices with `out` — while the GPU works on the blocks behind it."""
        from empanada_b200 import _cabi as C
        from empanada_b200.inference import postprocess as pp
        dev = packed.device
        n = len(zs)
        bad = np.zeros(n, dtype=bool)
        n_runs = np.zeros(n, dtype=np.int64)
        inst = [None] * n
        worst = 0
        SB, hw_, flags = subs['SB'], subs['host_words'], subs['flags']
        host_np = subs['host'].numpy()
        t_first = time.perf_counter()
        for bi in range(subs['n_sub']):
            i0 = bi * SB
            B = min(SB, n - i0)
            spins = 0
            while flags[bi] == 0:                                   # written by the copy engine after the block's tables
                spins += 1
                if spins % 4096 == 0 and _copy_stream(dev).query():  # the stream drained without the flag: a CUDA error
                    torch.cuda.synchronize(dev)
                    if flags[bi] == 0:
                        raise RuntimeError('emp_stack_blocks: block tables never arrived')
            if bi == 0:
                self._marks['first_block_arrival_s'] = time.perf_counter() - t_first
            words = host_np[bi * hw_:(bi + 1) * hw_]
            need = _BlockTables.words_needed(words, B)
            if need > hw_:                                          # the size guess was short: fetch the whole block (rare)
                words = packed[bi, :need].cpu().numpy()
            t = _BlockTables(words, B)
            worst = max(worst, (need - C.BLK_HDR_WORDS) // B + 1)
            for b in range(B):
                i = i0 + b
                fl = int(t.slices[b, 4])
                if fl & (C.FLAG_K_OVERFLOW | C.FLAG_RLE_OVERFLOW):
                    if streamed:                                    # the slice's heads are gone: nothing to redo it from
                        raise RuntimeError(f'slice {zs[i]} overflowed the deferred tables of a streamed stack; raise run_cap / '
                                           f'inst_cap, or finish() without advance()')
                    bad[i] = True
                    out._add_dict(zs[i], _slice_sync(self.engine, self.heads[zs[i]], sem8[i].view(1, 1, *self._plane), self.labels,
                                                     self.upsampling, self.force_connected))
                    continue
                pp._check_flags(fl)
                out._add(zs[i], t, b)
                n_runs[i] = int(t.slices[b, 3])
                inst[i] = out.inst_rows(zs[i])
        key = (dev.index,) + tuple(self._plane)
        _words_per_slice[key] = max(_words_per_slice.get(key, 0), worst)
        return bad, n_runs, inst

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

    def finish(self):
        """Post-process + RLE-encode this rank's block.  Returns an RleStack ({z: rle_seg}, dicts built on access)."""
        with _gc_paused():          # a cyclic collection over a previous block's ~10^5 dicts would stall the enqueue loop
            if self._stream:
                return self._finish_streamed()
            return self._finish()

    def _finish(self):
        from empanada_b200 import _cabi as C
        self._marks = {}
        self._geom = None
        e = self.engine
        zs = list(range(self.z0, self.z1))
        assert all(z in self.heads for z in range(self.z0, self.z_halo)), 'add() every slice of slices() first'
        planes = [_f32c(self.heads[z]['sem']) for z in range(self.z0, self.z_halo)]
        if not planes[0].is_cuda:
            raise RuntimeError('StackShard runs on CUDA tensors only (there is no CPU fallback)')
        dev = _same_planes(planes)
        Cn, H, W = planes[0].shape[1:]
        self._plane = (int(H), int(W))
        t_a = time.perf_counter()
        need = None
        if all(0 <= int(c) < 64 for c in e.thing_list):
            hh, ww, _ = self._geometry(H, W)
            need = _device_pool.acquire(dev, (len(zs), hh * ww), torch.uint8).zero_()
        sem8, changed = self._chain(planes, dev, H * W, Cn, need)
        t_chain = time.perf_counter()
        subs = self._enqueue_blocks(zs, sem8, dev, H, W, need)
        t_blocks = time.perf_counter()
        return self._complete(zs, dev, subs, sem8, changed, (t_a, t_chain, t_blocks), streamed=False)

    def _complete(self, zs, dev, subs, sem8, changed, t_marks, streamed):
        """The tail every finish shares: all-gather of the label maxima (+ the carry flag), collection of the sub-blocks'
        tables as they land, label offsets."""
        from empanada_b200 import _cabi as C
        e = self.engine
        t_a, t_chain, t_blocks = t_marks
        packed, runs_all, counts = subs['packed'], subs['runs_all'], subs['maxlab'][:subs['nl']]
        # instance counts per class (+ the "my outgoing carry moved" flag) over all ranks
        multi = self.world > 1 and dist.is_available() and dist.is_initialized()
        table_h = None
        if multi:
            mine = torch.cat([counts, (changed if changed is not None else torch.zeros((1,), dtype=torch.int32, device=dev)).to(torch.int64)])
            gathered = torch.empty((self.world, mine.numel()), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(gathered, mine, group=self.group)
            table_h = _pinned_words(dev, 'counts', gathered.numel())[:gathered.numel()].view(gathered.shape)
            table_h.copy_(gathered, non_blocking=True)
        t_b = time.perf_counter()
        out = RleStack(self.labels, e.thing_list, e.label_divisor)
        bad, n_runs, inst = self._collect(zs, subs, packed, sem8, out, streamed)
        self._marks['collect_s'] = time.perf_counter() - t_b
        torch.cuda.current_stream(dev).synchronize()
        t_c = time.perf_counter()
        offs = {c: 0 for c in self.labels}
        if multi:
            table = table_h.numpy()
            if table[:, -1].any() and streamed:
                # every rank sees the same table: all of them raise
                ranks = np.flatnonzero(table[:, -1]).tolist()
                raise RuntimeError(f'streamed stack: the median carry of rank(s) {ranks} did not settle inside their first '
                                   f'sub-block of {self.block} slices; use a larger `block`, or finish() without advance()')
            if table[:, -1].any() and not self._settle:
                # some rank's repair reached the end of its block: its neighbour's carry was not final.  Every rank
                # sees the same table, so all of them settle the carries with further rounds and redo the block.
                self._settle = True
                try:
                    return self._finish()
                finally:
                    self._settle = False
            per_class = table[:, :-1]
            if bad.any() or (per_class >= C.BLK_MAXLAB_OVERFLOW).any():
                # a slice redone synchronously is not in the device counts.  The overflowing rank poisoned its row of the
                # gathered table (EMP_BLK_MAXLAB_OVERFLOW), so every rank raises here and nobody is left in a collective.
                ranks = np.flatnonzero((per_class >= C.BLK_MAXLAB_OVERFLOW).any(1)).tolist()
                raise RuntimeError(f'a slice overflowed the deferred tables on rank(s) {ranks} of a multi-rank stack; '
                                   f'raise run_cap / inst_cap')
            before = per_class[:self.rank].sum(0)
            worst = per_class.sum(0)
            for c, b, w in zip(self.labels, before.tolist(), worst.tolist()):
                if c in e.thing_list:
                    if w >= e.label_divisor:                        # identical on every rank: all of them raise
                        raise ValueError(f'class {c}: {w} instance labels over {self.world} ranks leave the class range; '
                                         f'raise label_divisor (reference default for 3D is 20000)')
                    offs[c] = int(b)
        out.offsets = offs
        self.label_offsets_ = offs
        self.timing_ = {'enqueue_s': t_b - t_a, 'wait_s': t_c - t_b, 'enqueue_chain_s': t_chain - t_a, 'enqueue_blocks_s': t_blocks - t_chain,
                        'enqueue_gather_s': t_b - t_blocks, 'after_s': time.perf_counter() - t_c}
        self.timing_.update(self._marks)
        self.tables_shape_ = self._plane
        self.tables_ = {'runs_all': runs_all, 'n_runs': n_runs, 'bad': bad, 'inst': inst, 'zs': zs,
                        'slot_areas': [None if r is None else r[:, 8] for r in inst]}
        self._keep = None
        return out

    # ------------------------------------------------------------------------------------------------------
    # streaming: sub-blocks are chained and post-processed while later slices are still being produced
    # ------------------------------------------------------------------------------------------------------
    def advance(self):
        """Optional, between add() calls: run the median chain and the post-processing of every sub-block (`block`
        slices) whose slices and `median_kernel_size // 2` look-ahead slices have been add()ed, on the current stream,
        and drop this object's references to their head tensors — the reference's slice-by-slice loop
        (engines.py:363-394, pdl_inference3d.py:163-187) at sub-block granularity: the work hides behind the CNN of the
        following slices, and only one sub-block of heads (plus a byte per voxel of hardened classes) stays resident.
        finish() then only has the tail left.  Returns the number of sub-blocks enqueued by this call.

        On ranks > 0 of a multi-rank stack the first sub-block waits for the block's end: it is chained from the guessed
        carry at once (the following sub-blocks continue from its end state), but its tables are cut after the true carry
        has arrived and the sub-block has been repaired — the call that reaches the block's last sub-block exchanges the
        carries with the neighbour ranks (a collective: every rank must get there); a carry that does not settle inside
        that first sub-block raises on every rank at finish().  Not streamed (advance() returns 0 and finish() does everything): blocks of a single sub-block, and
        multi-class probabilities on a multi-rank stack (their carry is handed over rank after rank)."""
        with _gc_paused():
            return self._advance(final=False)

    def _on_side(self, dev):
        """Context of the streamed launches: the side stream (after it has caught up with the producer of the heads) or
        the current stream."""
        if self.side_stream is None:
            return contextlib.nullcontext()
        self.side_stream.wait_stream(torch.cuda.current_stream(dev))
        return torch.cuda.stream(self.side_stream)

    def _forget(self, zs):
        """Drop the references to the heads of slices whose launches are enqueued."""
        for z in zs:
            h = self.heads[z]
            if self.side_stream is not None:                        # the allocator must not reuse them before the side stream is done
                for k in ('sem', 'ctr_hmp', 'offsets'):
                    if isinstance(h.get(k), torch.Tensor) and h[k].is_cuda:
                        h[k].record_stream(self.side_stream)
            self.heads[z] = {'size': h['size']}

    def _stream_begin(self):
        from empanada_b200 import _cabi as C
        e = self.engine
        first = _f32c(self.heads[self.z0]['sem'])
        if not first.is_cuda:
            raise RuntimeError('StackShard runs on CUDA tensors only (there is no CPU fallback)')
        dev = C.require_cuda(first)
        assert first.dim() == 4 and first.size(0) == 1
        Cn, H, W = (int(v) for v in first.shape[1:])
        n, mid = self.z1 - self.z0, self.mid
        multi = self.world > 1 and mid > 0 and dist.is_available() and dist.is_initialized()
        if n <= self.block or (multi and Cn > 1):
            self._stream = False                                    # not streamable: finish() does everything
            return
        self._marks = {}
        self._geom = None
        self._plane = (H, W)
        hw = H * W
        zs = list(range(self.z0, self.z1))
        need = None
        if all(0 <= int(c) < 64 for c in e.thing_list):
            hh, ww, _ = self._geometry(H, W)
            need = _device_pool.acquire(dev, (n, hh * ww), torch.uint8).zero_()
        carry = {}
        if mid > 0:
            for name in ('out', 'a', 'b', 'p0', 'p1'):
                carry[name] = [torch.empty((Cn * hw,), dtype=torch.float32, device=dev) for _ in range(mid)]
            carry['guess'] = [first.reshape(-1).clone()] * mid      # this rank's first raw plane
        self._stream = {
            'dev': dev, 'Cn': Cn, 'H': H, 'W': W, 'hw': hw, 'multi': multi, 'zs': zs, 'need': need, 'carry': carry,
            'sem8': _device_pool.acquire(dev, (n, hw), torch.uint8),
            'best': torch.empty((n, hw), dtype=torch.float32, device=dev) if Cn > 1 else None,
            'prep': self._blocks_prepare(zs, dev, H, W), 'next': 0, 'state': None, 'keep': [], 'group0': None,
            't_a': time.perf_counter(),
        }

    def _carry_table(self, st, name):
        """Device array of the `mid` plane pointers of one carry buffer (what the chain kernels take)."""
        t = torch.tensor([p.data_ptr() for p in st['carry'][name]], dtype=torch.int64).to(st['dev'])
        st['keep'].append(t)
        return t.data_ptr()

    def _advance(self, final):
        from empanada_b200 import _cabi as C
        if self._stream is None:
            if self.z0 not in self.heads:
                return 0
            self._stream_begin()
        st = self._stream
        if st is False:
            return 0
        with self._on_side(st['dev']):
            return self._advance_groups(st, final)

    def _advance_groups(self, st, final):
        from empanada_b200 import _cabi as C
        e, L = self.engine, C.lib()
        dev, hw, Cn, mid, ks = st['dev'], st['hw'], st['Cn'], self.mid, self.ks
        prep = st['prep']
        SB, n, n_sub = prep['SB'], prep['n'], prep['n_sub']
        stream = C.stream_ptr(dev)
        vp = ctypes.c_void_p
        thr = float(e.confidence_thr)
        done = 0
        while st['next'] < n_sub:
            g = st['next']
            i0 = g * SB
            m = min(SB, n - i0)
            z_need = min(self.z0 + i0 + m + mid, self.z_halo)       # the group's slices + the chain's look-ahead
            if not all(z in self.heads for z in range(self.z0 + i0, z_need)):
                assert not final, 'add() every slice of slices() first'
                break
            planes = [_f32c(self.heads[z]['sem']) for z in range(self.z0 + i0, z_need)]
            assert _same_planes(planes) == dev and planes[0].shape == (1, Cn, st['H'], st['W'])
            tab = torch.tensor([p.data_ptr() for p in planes], dtype=torch.int64).to(dev)
            st['keep'].append((tab, planes if (g == 0 and st['multi'] and self.rank > 0) else None))
            # carry in: the stack's start (none), the guess (first group of a rank > 0), else the previous group's end state
            if g == 0:
                cin = self._carry_table(st, 'guess') if (st['multi'] and self.rank > 0) else None
            else:
                cin = st['state']
            cout = None
            if mid > 0:
                cout = self._carry_table(st, 'out' if g + 1 == n_sub else ('p0' if g % 2 == 0 else 'p1'))
            nm = None
            if st['need'] is not None:
                _, ww, shift = self._geometry(st['H'], st['W'])
                bits = 0
                for c in e.thing_list:
                    bits |= 1 << int(c)
                nm = C.NeedMap(map=st['need'].data_ptr() + i0 * st['need'].shape[1], stride=st['need'].shape[1], W=st['W'],
                               shift=shift, wc=ww, thing_bits=bits)
            with torch.cuda.device(dev):
                C.check(L.emp_median_chain(vp(tab.data_ptr()), m, len(planes), self.z0 + i0, self.depth, ks, Cn, hw,
                                           vp(cin) if cin else None, thr, vp(st['sem8'].data_ptr() + i0 * hw), hw,
                                           vp(st['best'].data_ptr() + 4 * i0 * hw) if st['best'] is not None else None,
                                           vp(cout) if cout else None, ctypes.byref(nm) if nm is not None else None, stream))
            st['state'] = cout
            if g + 1 == n_sub and st['multi']:
                # the block's end state exists: hand it to rank + 1, take rank - 1's, repair the first sub-block from it and
                # cut its tables — before this sub-block's own launches, so that the hand-over does not queue behind them
                self._stream_carry(st)
            if g == 0 and st['multi'] and self.rank > 0:
                st['group0'] = (tab, len(planes), m, nm)             # repaired and cut into tables once the true carry is here
            else:
                self._blocks_launch(prep, st['zs'], st['sem8'], st['need'], i0, m)
                self._forget(range(self.z0 + i0, self.z0 + i0 + m))
            st['next'] += 1
            done += 1
        return done

    def _stream_carry(self, st):
        from empanada_b200 import _cabi as C
        L = C.lib()
        dev, hw = st['dev'], st['hw']
        vp = ctypes.c_void_p
        exchange_planes(st['carry']['out'], st['carry']['a'], self.rank, self.world, self.group)
        if self.rank == 0:
            return
        tab, n_planes, m, nm = st['group0']
        st['changed'] = torch.zeros((1,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            C.check(L.emp_median_chain_repair(vp(tab.data_ptr()), m, n_planes, self.z0, self.depth, self.ks, hw,
                                              vp(self._carry_table(st, 'guess')), vp(self._carry_table(st, 'a')),
                                              float(self.engine.confidence_thr), vp(st['sem8'].data_ptr()), hw,
                                              vp(self._carry_table(st, 'b')), vp(st['changed'].data_ptr()),
                                              ctypes.byref(nm) if nm is not None else None, C.stream_ptr(dev)))
        self._blocks_launch(st['prep'], st['zs'], st['sem8'], st['need'], 0, m)
        self._forget(range(self.z0, self.z0 + m))

    def _finish_streamed(self):
        st = self._stream
        self._advance(final=True)
        if self.side_stream is not None:
            torch.cuda.current_stream(st['dev']).wait_stream(self.side_stream)
        t_blocks = time.perf_counter()
        self._keep = st                                              # alive until the stream has run
        self._stream = None
        try:
            return self._complete(st['zs'], st['dev'], st['prep'], st['sem8'], st.get('changed'), (st['t_a'], t_blocks, t_blocks),
                                  streamed=True)
        finally:
            self._keep = None
