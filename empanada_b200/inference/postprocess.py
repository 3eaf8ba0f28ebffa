"""
Drop-in for ``empanada.inference.postprocess`` (reference empanada/inference/postprocess.py):
same names, positional order, defaults, return shapes / dtypes and exceptions — every function
runs as hand-written sm_100a kernels through the C ABI of libempanada_b200.so.

Differences a caller can observe: tensors must live on a CUDA device (no CPU fallback — a CPU
tensor raises RuntimeError); semantic class ids must lie in [0, 4096); at most 16 thing classes;
``merge_semantic_and_instance`` needs non-negative instance ids.
"""
import ctypes
from typing import List

import torch
import torch.nn.functional as F

from empanada_b200 import _cabi as C

__all__ = [
    'factor_pad',
    'find_instance_center',
    'group_pixels',
    'get_instance_segmentation',
    'get_panoptic_segmentation'
]

DEFAULT_K_CAP = 32768      # centers per tile before the (rare) retry with a full-size table


def factor_pad(tensor, factor: int = 16):
    """Zero-pad bottom/right so H and W are divisible by ``factor`` (postprocess.py:25-36)."""
    h, w = tensor.size()[2:]
    pad_bottom = (factor - h % factor) % factor
    pad_right = (factor - w % factor) % factor
    if pad_bottom == 0 and pad_right == 0:
        return tensor
    return F.pad(tensor, (0, pad_right, 0, pad_bottom))


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _i64c(t):
    return t.detach().to(torch.int64).contiguous()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _check_flags(flags):
    if flags & C.FLAG_CLASS_RANGE:
        raise ValueError(f'semantic class ids must lie in [0, {C.MAX_CLASSES})')
    if flags & C.FLAG_ID_RANGE:
        raise ValueError('instance ids out of the supported range')


def find_instance_center(ctr_hmp, threshold: float = 0.1, nms_kernel: int = 7):
    """Center points of the heat-map: threshold -> k x k max-pool NMS -> all peaks in row-major
    order (postprocess.py:38-76).  ctr_hmp (N,1,H,W) with N == 1 -> (K,2) int64 (y,x)."""
    dev = C.require_cuda(ctr_hmp)
    hm = _f32c(ctr_hmp).squeeze()
    assert len(hm.size()) == 2, 'Something is wrong with center heatmap dimension.'
    H, W = hm.shape
    L = C.lib()
    cap = 4096
    while True:
        nbytes = L.emp_workspace_bytes(H, W, cap, 1)
        ws = C.workspace(dev, nbytes)
        out = torch.empty((cap, 2), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            C.check(L.emp_find_centers(_ptr(hm), H, W, float(threshold), int(nms_kernel), _ptr(out), cap,
                                       _ptr(ws), ws.numel(), C.stream_ptr(dev)))
        K = int(C.read_status(ws)[C.ST_K])
        if K <= cap:
            return out[:K]
        cap = K


def group_pixels(ctr, offsets, chunksize: int = 20, step: float = 1):
    """Instance id (1-based index of the nearest center to pixel + offset) for every pixel
    (postprocess.py:118-169).  ctr (K,2), offsets (1,2,H,W) -> (1,H,W) int64."""
    assert ctr.size(0) > 0
    if offsets.size(0) != 1:
        raise ValueError('Only supports inference for batch size = 1')
    dev = C.require_cuda(ctr, offsets)
    off = _f32c(offsets).squeeze(0)
    H, W = off.shape[1:]
    ids = torch.empty((1, H, W), dtype=torch.int64, device=dev)
    _group_pixels_into(ctr, off, H, W, chunksize, step, ids, False)
    return ids


def _group_pixels_into(ctr, off, H, W, chunksize, step, ids, ids_i32):
    dev = off.device
    ctr = _i64c(ctr)
    K = ctr.size(0)
    L = C.lib()
    nbytes = L.emp_workspace_bytes(H, W, K, 1)
    ws = C.workspace(dev, nbytes)
    with torch.cuda.device(dev):
        C.check(L.emp_group_pixels(_ptr(ctr), K, _ptr(off), H, W, float(step), int(chunksize), _ptr(ids),
                                   int(ids_i32), _ptr(ws), ws.numel(), C.stream_ptr(dev)))


def get_instance_segmentation(sem_seg, ctr_hmp, offsets, thing_list: List[int],
                              threshold: float = 0.1, nms_kernel: int = 7):
    """Class-agnostic instance ids on thing pixels (postprocess.py:171-221).
    Returns (thing_seg (1,H,W) int64, ctr (1,K,2) int64)."""
    assert sem_seg.size(0) == 1, 'Only batch size of 1 is supported!'
    if offsets.size(0) != 1:
        raise ValueError('Only supports inference for batch size = 1')
    dev = C.require_cuda(sem_seg, ctr_hmp, offsets)
    sem = _i64c(sem_seg)[0]
    hm = _f32c(ctr_hmp).squeeze()
    assert len(hm.size()) == 2, 'Something is wrong with center heatmap dimension.'
    off = _f32c(offsets).squeeze(0)
    H, W = hm.shape
    things, nt = C.i64_array(thing_list)
    L = C.lib()
    k_cap = min(DEFAULT_K_CAP, H * W)
    while True:
        nbytes = L.emp_workspace_bytes(H, W, k_cap, max(nt, 1))
        ws = C.workspace(dev, nbytes)
        ins = torch.empty_like(sem)
        ctr = torch.empty((k_cap, 2), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            C.check(L.emp_instance_segmentation(_ptr(sem), _ptr(hm), _ptr(off), H, W, things, nt,
                                                float(threshold), int(nms_kernel), _ptr(ins), _ptr(ctr),
                                                k_cap, k_cap, _ptr(ws), ws.numel(), C.stream_ptr(dev)))
        st = C.read_status(ws)
        K = int(st[C.ST_K])
        if K <= k_cap:
            break
        k_cap = K
    _check_flags(int(st[C.ST_FLAGS]))
    return ins, ctr[:K].unsqueeze(0)


def merge_semantic_and_instance(sem_seg, ins_seg, label_divisor: int, thing_list: List[int],
                                stuff_area: int, void_label: int):
    """Majority-vote merge of semantic and instance maps (postprocess.py:223-296).
    Output shape is the broadcast of the two inputs, dtype int64."""
    dev = C.require_cuda(sem_seg, ins_seg)
    out_shape = torch.broadcast_shapes(sem_seg.shape, ins_seg.shape)
    H, W = out_shape[-2:]
    if sem_seg.numel() != H * W or ins_seg.numel() != H * W:
        raise NotImplementedError('merge_semantic_and_instance: only single-image inputs '
                                  f'(got {tuple(sem_seg.shape)} and {tuple(ins_seg.shape)})')
    sem = _i64c(sem_seg).reshape(H, W)
    ins = _i64c(ins_seg).reshape(H, W)
    lo, hi = (int(v) for v in torch.aminmax(ins))
    if lo < 0:
        raise NotImplementedError('negative instance ids are not supported')
    things, nt = C.i64_array(thing_list)
    L = C.lib()
    nbytes = L.emp_workspace_bytes(H, W, hi, max(nt, 1))
    ws = C.workspace(dev, nbytes)
    pan = torch.empty(out_shape, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        C.check(L.emp_merge(_ptr(sem), _ptr(ins), H, W, int(label_divisor), things, nt, int(stuff_area),
                            int(void_label), hi, _ptr(pan), _ptr(ws), ws.numel(), C.stream_ptr(dev)))
    _check_flags(int(C.read_status(ws)[C.ST_FLAGS]))
    return pan


def _panoptic_tiles(sem, hm, off, thing_list, label_divisor, stuff_area, void_label, threshold,
                    nms_kernel, want_centers=True, k_cap=None, check=True):
    """Fused get_panoptic_segmentation over B tiles.  sem (B,H,W) int64|uint8, hm (B,H,W) f32,
    off (B,2,H,W) f32, all contiguous CUDA tensors.  Returns (pan (B,H,W) int64, centers (B,cap,2) or
    None, K per tile (list) or None).  With check=False nothing is read back (no sync)."""
    dev = sem.device
    B, H, W = sem.shape
    things, nt = C.i64_array(thing_list)
    L = C.lib()
    k_cap = min(k_cap or DEFAULT_K_CAP, H * W)
    sem_u8 = int(sem.dtype == torch.uint8)
    while True:
        per_tile = L.emp_workspace_bytes(H, W, k_cap, max(nt, 1))
        ws = C.workspace(dev, per_tile * B)
        pan = torch.empty((B, H, W), dtype=torch.int64, device=dev)
        ctr = torch.empty((B, k_cap, 2), dtype=torch.int64, device=dev) if want_centers else None
        with torch.cuda.device(dev):
            C.check(L.emp_panoptic_batched(B, _ptr(sem), sem_u8, _ptr(hm), _ptr(off), H, W, things, nt,
                                           int(label_divisor), int(stuff_area), int(void_label),
                                           float(threshold), int(nms_kernel), _ptr(pan),
                                           _ptr(ctr) if want_centers else None, k_cap if want_centers else 0,
                                           k_cap, _ptr(ws), per_tile, C.stream_ptr(dev)))
        if not check:
            return pan, ctr, None
        st = C.read_status(ws, B, per_tile).reshape(B, -1)
        Ks = [int(v) for v in st[:, C.ST_K]]
        if max(Ks) <= k_cap:
            break
        k_cap = max(Ks)
    for f in st[:, C.ST_FLAGS]:
        _check_flags(int(f))
    return pan, ctr, Ks


def get_panoptic_segmentation(sem, ctr_hmp, offsets, thing_list: List[int], label_divisor: int,
                              stuff_area: int, void_label: int, threshold: float = 0.1,
                              nms_kernel: int = 7):
    """Panoptic post-processing (postprocess.py:298-356).  sem (1,1,H,W) hardened classes,
    ctr_hmp (1,1,H,W), offsets (1,2,H,W) -> (pan_seg (1,1,H,W) int64, center (1,K,2) int64)."""
    if sem.size(1) != 1:
        raise ValueError('Expect single channel semantic segmentation. Softmax/argmax first!')
    if sem.size(0) != 1:
        raise ValueError('Only supports inference for batch size = 1')
    if ctr_hmp.size(0) != 1:
        raise ValueError('Only supports inference for batch size = 1')
    if offsets.size(0) != 1:
        raise ValueError('Only supports inference for batch size = 1')
    C.require_cuda(sem, ctr_hmp, offsets)
    H, W = sem.shape[-2:]
    sem_t = sem.detach() if sem.dtype == torch.uint8 else _i64c(sem)
    sem_t = sem_t.contiguous().reshape(1, H, W)
    hm = _f32c(ctr_hmp).reshape(1, H, W)
    off = _f32c(offsets).reshape(1, 2, H, W)
    pan, ctr, Ks = _panoptic_tiles(sem_t, hm, off, thing_list, label_divisor, stuff_area, void_label,
                                   threshold, nms_kernel)
    return pan.reshape(1, 1, H, W), ctr[:, :Ks[0]]
