"""
Drop-in for ``empanada.inference.rle`` (reference empanada/inference/rle.py): the dense panoptic
map -> nested run-length dict conversion runs on the GPU (libempanada_b200 ``emp_rle``: row-run
extraction, run-based 8-connected components, raster-order renumbering, run merging across row
ends, boxes); only the runs (KBs) cross to the host, where the dict is assembled.
"""
import ctypes

import numpy as np
import torch

from empanada_b200 import _cabi as C

__all__ = [
    'pan_seg_to_rle_seg',
    'rle_seg_to_pan_seg',
    'unpack_rle_attrs'
]


def _as_cuda_i64(pan_seg, device=None):
    if isinstance(pan_seg, np.ndarray):
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        t = torch.from_numpy(np.ascontiguousarray(pan_seg).astype(np.int64, copy=False)).to(device)
    else:
        t = pan_seg.detach()
        C.require_cuda(t)
        t = t.to(torch.int64)
    t = t.squeeze()
    if t.dim() != 2:
        raise ValueError(f'pan_seg must be (h, w); got {tuple(pan_seg.shape)}')
    return t.contiguous()


def rle_enqueue(pan, labels, label_divisor, thing_list, force_connected, runs_out, inst_out):
    """Enqueue emp_rle for a contiguous (H,W) int64 CUDA tensor on the current stream, writing into
    the caller's runs_out (run_cap,3) / inst_out (inst_cap,8) int64 CUDA tensors.  No host
    synchronisation.  Returns the workspace tensor whose first 64 bytes are the status block
    (row-runs / runs / instances found, overflow flag) once the stream has run."""
    dev = pan.device
    H, W = pan.shape
    L = C.lib()
    labels_a, nl = C.i64_array(labels)
    things_a, nt = C.i64_array(thing_list)
    run_cap, inst_cap = runs_out.shape[0], inst_out.shape[0]
    nbytes = L.emp_rle_workspace_bytes(H, W, run_cap, nl, int(label_divisor))
    if nbytes == 0:
        raise ValueError('bad arguments to pan_seg_to_rle_seg')
    ws = C.workspace(dev, nbytes, 'rle')
    with torch.cuda.device(dev):
        C.check(L.emp_rle(ctypes.c_void_p(pan.data_ptr()), H, W, labels_a, nl, int(label_divisor), things_a, nt,
                          int(bool(force_connected)), ctypes.c_void_p(runs_out.data_ptr()), run_cap,
                          ctypes.c_void_p(inst_out.data_ptr()), inst_cap, ctypes.c_void_p(ws.data_ptr()),
                          ws.numel(), C.stream_ptr(dev)))
    return ws


def rle_tables(pan_seg, labels, label_divisor, thing_list, force_connected=True, run_cap=None,
               device=None):
    """GPU part: returns host numpy tables (inst (n_inst, 8) int64, runs (n_runs, 3) int64) as
    described in include/empanada_b200.h (runs in ascending start order, column 2 = slot)."""
    pan = _as_cuda_i64(pan_seg, device)
    dev = pan.device
    H, W = pan.shape
    if run_cap is None:
        run_cap = max(1 << 16, (H * W) // 64)
    inst_cap = run_cap
    while True:
        runs = torch.empty((run_cap, 3), dtype=torch.int64, device=dev)
        inst = torch.empty((inst_cap, 8), dtype=torch.int64, device=dev)
        ws = rle_enqueue(pan, labels, label_divisor, thing_list, force_connected, runs, inst)
        st = C.read_status(ws)
        n_rowruns, n_runs, n_inst = int(st[C.ST_NROWRUNS]), int(st[C.ST_NRUNS]), int(st[C.ST_NINST])
        if not (int(st[C.ST_FLAGS]) & C.FLAG_RLE_OVERFLOW):
            break
        run_cap = max(n_rowruns, n_runs, n_inst, run_cap * 2)
        inst_cap = run_cap
    return inst[:n_inst].cpu().numpy(), runs[:n_runs].cpu().numpy()


def grouped_to_rle_seg(inst, starts, lens, labels):
    """Assemble the reference's nested dict {class: {label: {'box', 'starts', 'runs'}}} from the
    instance table and the runs already grouped by instance slot (ascending start inside a slot).
    The per-instance arrays are views into `starts` / `lens` (no per-instance copies)."""
    rle_seg = {int(l): {} for l in labels}
    n = inst.shape[0]
    if n == 0:
        return rle_seg
    b = np.concatenate(([0], np.cumsum(inst[:, 6]))).tolist()
    boxes = inst[:, 2:6].tolist()
    cls, labs = inst[:, 0].tolist(), inst[:, 1].tolist()
    for i in range(n):
        rle_seg[cls[i]][labs[i]] = {'box': tuple(boxes[i]), 'starts': starts[b[i]:b[i + 1]], 'runs': lens[b[i]:b[i + 1]]}
    return rle_seg


def tables_to_rle_seg(inst, runs, labels):
    """Host part: group the start-ordered runs by instance slot (stable) and assemble the dict."""
    if inst.shape[0] == 0:
        return {int(l): {} for l in labels}
    # runs are in ascending start order, so a plain sort of (slot, position) is the stable grouping
    order = np.argsort(runs[:, 2] * (1 << 32) + np.arange(runs.shape[0], dtype=np.int64))
    return grouped_to_rle_seg(inst, runs[order, 0], runs[order, 1], labels)


def pan_seg_to_rle_seg(pan_seg, labels, label_divisor, thing_list, force_connected=True):
    r"""Converts a panoptic segmentation to run length encodings (rle.py:26-86).

    pan_seg may be a numpy array (as in the reference) or — to skip the D2H/H2D round trip — the
    CUDA tensor an engine returned.  Returns {class: {instance label: {'box', 'starts', 'runs'}}}.
    """
    inst, runs = rle_tables(pan_seg, labels, label_divisor, thing_list, force_connected)
    return tables_to_rle_seg(inst, runs, labels)


def rle_seg_to_pan_seg(rle_seg, shape):
    r"""Converts run length encodings to a panoptic segmentation (rle.py:88-118).  Host numpy,
    as in the reference (it is only used by tests and by consumers of the host tracker)."""
    pan_seg = np.zeros(shape, dtype=np.uint32).ravel()
    for instance_attrs in rle_seg.values():
        for object_id, attrs in instance_attrs.items():
            for s, r in zip(attrs['starts'], attrs['runs']):
                pan_seg[s:s + r] = object_id
    return pan_seg.reshape(shape)


def _string_to_rle(encoding):
    enc = np.array([int(i) for i in encoding.split(' ')])
    return enc[::2], enc[1::2]


def unpack_rle_attrs(instance_rle_seg):
    r"""Unpacks one class's rle dict into (labels, boxes, starts list, runs list) (rle.py:120-150)."""
    labels, boxes, starts, runs = [], [], [], []
    for label, attrs in instance_rle_seg.items():
        labels.append(int(label))
        boxes.append(attrs['box'])
        if 'rle' in attrs:
            s, r = _string_to_rle(attrs['rle'])
            starts.append(s)
            runs.append(r)
        else:
            starts.append(attrs['starts'])
            runs.append(attrs['runs'])
    return np.array(labels), np.array(boxes), starts, runs
