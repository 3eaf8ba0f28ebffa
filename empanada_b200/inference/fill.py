"""
Dense fill of run-length encoded instances on the GPU — the role of
``empanada.array_utils.numpy_fill_instances`` (reference array_utils.py:725-736; used by the
inference scripts to write the labelled volume): every run ``[start, start + length)`` of an instance
is painted with the instance's label into a flat view of the volume; voxels no run covers keep their
value.  libempanada_b200 ``emp_fill_runs`` does it with one warp per run.

  * ``fill_instances(volume, instances)`` — the reference's call shape, for a CUDA tensor volume and a
    dict {label: {'starts', 'runs', ...}} (runs index the flattened volume);
  * ``fill_block(runs_all, n_runs, labels, shape, dtype)`` — for the z-sharded stack driver: the run
    tables of a whole z-block are already in HBM, `labels[s][slot]` is each instance's final label.
"""
import ctypes

import numpy as np
import torch

from empanada_b200 import _cabi as C

__all__ = ['fill_instances', 'fill_slabs', 'fill_block']


def _elem_bytes(t):
    if t.dtype in (torch.int64,):
        return 8
    if t.dtype in (torch.int32, torch.uint32):
        return 4
    raise TypeError(f'volume must be int64, int32 or uint32 (got {t.dtype})')


def _launch(runs, n_runs_dev, max_runs, labels, out2d, device):
    n_slices, run_stride = int(runs.shape[0]), int(runs.shape[1])
    with torch.cuda.device(device):
        C.check(C.lib().emp_fill_runs(ctypes.c_void_p(runs.data_ptr()), run_stride, ctypes.c_void_p(n_runs_dev.data_ptr()),
                                      n_slices, int(max_runs), ctypes.c_void_p(labels.data_ptr()), int(labels.shape[1]),
                                      ctypes.c_void_p(out2d.data_ptr()), _elem_bytes(out2d), int(out2d.shape[1]),
                                      C.stream_ptr(device)))


def _run_table(instances):
    """(table (n, 3) int64 rows (start, length, slot), labels list) of a dict {label: {'starts', 'runs'}}; slots number
    the instances in dict order."""
    starts = [np.asarray(a['starts'], np.int64) for a in instances.values()]
    lens = [np.asarray(a['runs'], np.int64) for a in instances.values()]
    slot = np.repeat(np.arange(len(starts), dtype=np.int64), [len(s) for s in starts])
    table = np.stack([np.concatenate(starts), np.concatenate(lens), slot], 1) if slot.size else np.zeros((0, 3), np.int64)
    return table, [int(lab) for lab in instances.keys()]


def _paint(flat, table, labels, dev):
    """Paint the runs of `table` (flat indices into the (1, n) CUDA tensor `flat`) with labels[slot].  Runs outside the
    tensor are clipped (the kernel clamps too).  Overlapping instances are painted in slot order, one launch per
    instance, so that later ones win as in the reference's loop."""
    if table.shape[0] == 0:
        return
    lab = torch.tensor([labels], dtype=torch.int64, device=dev)
    s_sorted = np.argsort(table[:, 0], kind='stable')
    ends = table[s_sorted, 0] + table[s_sorted, 1]
    overlapping = bool((table[s_sorted, 0][1:] < np.maximum.accumulate(ends)[:-1]).any())
    if not overlapping:
        groups = [table]
    else:                                                       # one group per instance, in slot order (one sort, not a mask per instance)
        order = np.argsort(table[:, 2], kind='stable')
        by_slot = table[order]
        cuts = np.flatnonzero(np.diff(by_slot[:, 2])) + 1
        groups = np.split(by_slot, cuts)
    for g in groups:
        if g.shape[0] == 0:
            continue
        runs = torch.from_numpy(np.ascontiguousarray(g[None])).to(dev)
        n_runs = torch.tensor([g.shape[0]], dtype=torch.int32, device=dev)
        _launch(runs, n_runs, g.shape[0], lab, flat, dev)


def fill_instances(volume, instances):
    """Fill a CUDA tensor `volume` (any shape, int64 / int32 / uint32, contiguous) in place with run
    length encoded instances {label: {'starts': ..., 'runs': ...}} and return it.  Later instances
    overwrite earlier ones where their runs overlap, as in the reference's loop."""
    dev = C.require_cuda(volume)
    assert volume.is_contiguous()
    if len(instances) == 0:
        return volume
    table, labels = _run_table(instances)
    _paint(volume.view(1, -1), table, labels, dev)
    return volume


def fill_slabs(volume, instances, max_slab_bytes=1 << 30, device=None):
    """Fill a HOST-side (d, h, w) volume in place — a numpy array or anything sliceable along z the way a zarr array
    is (``volume[z0:z1]`` reads / writes a slab; ``.chunks[0]`` gives the preferred slab height) — one z-slab at a
    time: the slab's part of every run is painted in HBM (emp_fill_runs) and written back.  Only one slab is ever
    resident, so volumes larger than the GPU's (or the host's) memory work, as with the reference's chunk-wise zarr
    fill (zarr_utils.py:88-175)."""
    if len(instances) == 0:
        return volume
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    d, h, w = (int(v) for v in volume.shape)
    plane = h * w
    itemsize = np.dtype(volume.dtype).itemsize
    dz = max(1, int(max_slab_bytes // max(plane * max(itemsize, 4), 1)))
    zc = int(volume.chunks[0]) if hasattr(volume, 'chunks') else 1
    dz = max(dz - dz % zc, zc)                                     # whole chunks per slab: every chunk is written once
    dz = max(1, min(dz, d))
    table, labels = _run_table(instances)
    if table.shape[0] == 0:
        return volume
    s_all, e_all = table[:, 0], table[:, 0] + table[:, 1]
    for z0 in range(0, d, dz):
        z1 = min(d, z0 + dz)
        a, b = z0 * plane, z1 * plane
        keep = (e_all > a) & (s_all < b)
        if not keep.any():
            continue
        s_c = np.maximum(s_all[keep], a) - a
        e_c = np.minimum(e_all[keep], b) - a
        sub = np.stack([s_c, e_c - s_c, table[keep, 2]], 1)
        slab = np.ascontiguousarray(volume[z0:z1])
        kind = slab.dtype
        if kind.kind not in 'iu':
            raise Exception(f'Unsupported volume dtype {kind}')
        if kind.itemsize >= 4:
            carrier = np.int32 if kind.itemsize == 4 else np.int64
            t = torch.from_numpy(slab.view(carrier)).to(device)
            _paint(t.view(1, -1), sub, labels, device)
            volume[z0:z1] = t.cpu().numpy().view(kind)
        else:                                                       # narrow volumes are painted as int32 and narrowed back
            t = torch.from_numpy(slab.astype(np.int32)).to(device)
            _paint(t.view(1, -1), sub, labels, device)
            volume[z0:z1] = t.cpu().numpy().astype(kind)
    return volume


def fill_block(runs_all, n_runs, labels, shape, dtype=torch.int64, out=None):
    """Labelled planes of a z-block from its run tables: runs_all (n, run_stride, 3) int64 CUDA tensor,
    n_runs (n,) host counts, labels (n, max_slots) int64 (host array or CUDA tensor; negative = skip),
    shape (H, W) of a plane.  Returns an (n, H, W) tensor of `dtype` (zeros where nothing was painted)."""
    dev = runs_all.device
    n = int(runs_all.shape[0])
    H, W = shape
    if out is None:
        out = torch.zeros((n, H, W), dtype=dtype, device=dev)
    n_runs = np.asarray(n_runs, np.int64)
    lab = labels if torch.is_tensor(labels) else torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int64))
    lab = lab.to(dev).contiguous()
    nr = torch.from_numpy(n_runs.astype(np.int32)).to(dev)
    _launch(runs_all, nr, int(n_runs.max()) if n else 0, lab, out.view(n, H * W), dev)
    return out
