"""
oracle — CPU restatement of empanada's panoptic post-processing, median queue and RLE codec.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only as the checker
(or the timed CPU baseline).  Nothing in ``empanada_b200/`` imports it; the product path
raises if its CUDA library is missing instead of falling back here.

Parity: pinned — ``tests/golden/make_golden.py`` runs the *reference itself* (imported from
``/root/reference`` in the build container, ``rle.py`` through the skimage shim in
``oracle/ref_shim.py``) and stores inputs + outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against them bit for bit.

Two layers:
  * ``oracle.c`` (compiled to ``oracle/_build/liboracle.so`` by ``oracle.build()``): the closed
    forms of SURVEY.md Appendix A in plain C — fast enough for 4096x4096 checks.
  * ``oracle/torch_port.py``: the reference's *op sequence* (chunks of 20 centers, per-instance
    loop) restated with torch CPU ops — the CPU baseline timed by ``bench.py``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """gcc oracle.c -> oracle/_build/liboracle.so (idempotent)."""
    src = os.path.join(_HERE, "oracle.c")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(src)):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp",
           src, "-o", _SO, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, i32, f32, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p
        L.orc_find_centers.restype = i64
        L.orc_find_centers.argtypes = [vp, i32, i32, f32, i32, vp, i64]
        L.orc_group_pixels.restype = None
        L.orc_group_pixels.argtypes = [vp, i64, vp, i32, i32, f32, i32, vp]
        L.orc_group_pixels_masked.restype = None
        L.orc_group_pixels_masked.argtypes = [vp, i64, vp, i32, i32, f32, i32, vp, vp]
        L.orc_merge.restype = i32
        L.orc_merge.argtypes = [vp, vp, i64, i64, vp, i32, i64, i64, vp]
        L.orc_median.restype = None
        L.orc_median.argtypes = [vp, i32, i64, vp]
        L.orc_harden.restype = None
        L.orc_harden.argtypes = [vp, i32, i64, f32, vp]
        L.orc_ccl8.restype = i64
        L.orc_ccl8.argtypes = [vp, i32, i32, vp]
        L.orc_pan_to_rle.restype = i32
        L.orc_pan_to_rle.argtypes = [vp, i32, i32, vp, i32, i64, vp, i32, i32,
                                     vp, i64, vp, i64, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ---------------------------------------------------------------------------------------------
# postprocess.py restatements
# ---------------------------------------------------------------------------------------------
def find_instance_center(ctr_hmp, threshold=0.1, nms_kernel=7):
    """postprocess.py:38-76.  ctr_hmp: any array squeezable to (H,W) fp32 -> (K,2) int64."""
    hm = _c(np.squeeze(np.asarray(ctr_hmp)), np.float32)
    assert hm.ndim == 2
    H, W = hm.shape
    cap = 1024
    while True:
        out = np.empty((cap, 2), np.int64)
        k = lib().orc_find_centers(_p(hm), H, W, np.float32(threshold), int(nms_kernel), _p(out), cap)
        if k <= cap:
            return out[:k].copy()
        cap = int(k)


def group_pixels(ctr, offsets, chunksize=20, step=1.0):
    """postprocess.py:118-169.  ctr (K,2) int64, offsets (1,2,H,W)|(2,H,W) fp32 -> (1,H,W) int64."""
    ctr = _c(ctr, np.int64).reshape(-1, 2)
    assert ctr.shape[0] > 0
    off = _c(offsets, np.float32)
    if off.ndim == 4:
        if off.shape[0] != 1:
            raise ValueError('Only supports inference for batch size = 1')
        off = off[0]
    _, H, W = off.shape
    off = np.ascontiguousarray(off)
    ids = np.empty((H, W), np.int64)
    lib().orc_group_pixels(_p(ctr), ctr.shape[0], _p(off), H, W, np.float32(step), int(chunksize), _p(ids))
    return ids[None]


def get_instance_segmentation(sem_seg, ctr_hmp, offsets, thing_list, threshold=0.1, nms_kernel=7):
    """postprocess.py:171-221.  sem_seg (1,1,H,W) int64 -> ((1,H,W) int64, (1,K,2) int64)."""
    sem = _c(sem_seg, np.int64)
    assert sem.shape[0] == 1
    sem = sem[0]
    thing = np.ascontiguousarray(np.isin(sem, np.asarray(list(thing_list), np.int64)).astype(np.uint8))
    ctr = find_instance_center(ctr_hmp, threshold, nms_kernel)
    if ctr.shape[0] == 0:
        return np.zeros_like(sem), ctr[None]
    off = _c(offsets, np.float32)
    if off.ndim == 4:
        off = np.ascontiguousarray(off[0])
    _, H, W = off.shape
    ids = np.empty((1, H, W), np.int64)
    # search only thing pixels: the others are multiplied by 0 anyway (postprocess.py:221)
    lib().orc_group_pixels_masked(_p(ctr), ctr.shape[0], _p(off), H, W, np.float32(1.0), 20, _p(thing), _p(ids))
    return ids, ctr[None]


def merge_semantic_and_instance(sem_seg, ins_seg, label_divisor, thing_list, stuff_area, void_label):
    """postprocess.py:223-296.  Output shape = broadcast(sem_seg, ins_seg)."""
    sem = np.asarray(sem_seg, np.int64)
    ins = np.asarray(ins_seg, np.int64)
    shape = np.broadcast_shapes(sem.shape, ins.shape)
    sem_b = _c(np.broadcast_to(sem, shape), np.int64)
    ins_b = _c(np.broadcast_to(ins, shape), np.int64)
    things = _c(list(thing_list), np.int64)
    pan = np.empty(shape, np.int64)
    rc = lib().orc_merge(_p(sem_b), _p(ins_b), sem_b.size, int(label_divisor), _p(things),
                         things.size, int(stuff_area), int(void_label), _p(pan))
    assert rc == 0
    return pan


def get_panoptic_segmentation(sem, ctr_hmp, offsets, thing_list, label_divisor, stuff_area,
                              void_label, threshold=0.1, nms_kernel=7):
    """postprocess.py:298-356.  Returns (pan (1,1,H,W) int64, centers (1,K,2) int64)."""
    sem = np.asarray(sem)
    if sem.shape[1] != 1:
        raise ValueError('Expect single channel semantic segmentation. Softmax/argmax first!')
    for t in (sem, np.asarray(ctr_hmp), np.asarray(offsets)):
        if t.shape[0] != 1:
            raise ValueError('Only supports inference for batch size = 1')
    ins, ctr = get_instance_segmentation(sem, ctr_hmp, offsets, thing_list, threshold, nms_kernel)
    pan = merge_semantic_and_instance(sem, ins, label_divisor, thing_list, stuff_area, void_label)
    return pan, ctr


# ---------------------------------------------------------------------------------------------
# engines.py restatements
# ---------------------------------------------------------------------------------------------
def median_planes(planes):
    """engines.py:59-66: middle order statistic over an odd number of equal-shape fp32 planes."""
    planes = [_c(p, np.float32) for p in planes]
    ks = len(planes)
    assert ks % 2 == 1
    ptrs = (ctypes.c_void_p * ks)(*[p.ctypes.data for p in planes])
    out = np.empty_like(planes[0])
    lib().orc_median(ptrs, ks, planes[0].size, _p(out))
    return out


def harden_seg(prob, confidence_thr):
    """engines.py:114-121.  prob (1,C,H,W) fp32 -> (1,1,H,W) int64."""
    prob = _c(prob, np.float32)
    _, C, H, W = prob.shape
    out = np.empty((1, 1, H, W), np.int64)
    lib().orc_harden(_p(prob), C, H * W, np.float32(confidence_thr), _p(out))
    return out


class MedianQueue:
    """engines.py:47-90 restated on numpy planes, recursion included: ``get_next`` stores the
    median back into the queued entry, so later windows see filtered planes."""

    def __init__(self, median_kernel_size):
        assert median_kernel_size % 2 == 1, "Kernel size must be odd integer!"
        self.ks = median_kernel_size
        self.mid_idx = (median_kernel_size - 1) // 2
        self.q = []

    def enqueue(self, item):
        self.q.append(item)
        if len(self.q) > self.ks:
            self.q.pop(0)

    def get_next(self, keys):
        nq = len(self.q)
        if nq <= self.mid_idx:
            return self.q[-1]
        if nq < self.ks:
            return None
        out = self.q[self.mid_idx]
        for k in keys:
            out[k] = median_planes([e[k] for e in self.q])
        return out

    def end(self):
        return self.q[self.mid_idx + 1:]


def nearest_upsample(ids, scale):
    """F.interpolate(mode='nearest', scale_factor=int) for power-of-two scales (engines.py:274):
    out[Y, X] = in[Y // s, X // s]."""
    return np.repeat(np.repeat(ids, scale, axis=-2), scale, axis=-1)


# ---------------------------------------------------------------------------------------------
# rle.py restatements
# ---------------------------------------------------------------------------------------------
def connected_components(seg):
    """rle.py:18-24: 8-connected multi-value CCL numbered by raster-first pixel."""
    seg = _c(seg, np.int64)
    H, W = seg.shape
    out = np.empty((H, W), np.int64)
    lib().orc_ccl8(_p(seg), H, W, _p(out))
    return out


def pan_seg_to_rle_seg(pan_seg, labels, label_divisor, thing_list, force_connected=True):
    """rle.py:26-86 -> the same nested dict {class: {label: {'box','starts','runs'}}}."""
    pan = _c(pan_seg, np.int64)
    H, W = pan.shape
    labels_a = _c(list(labels), np.int64)
    things = _c(list(thing_list), np.int64)
    inst_cap, runs_cap = 1 << 12, 1 << 16
    while True:
        inst = np.empty((inst_cap, 7), np.int64)
        runs = np.empty((runs_cap, 2), np.int64)
        ni = ctypes.c_int64(0)
        nr = ctypes.c_int64(0)
        rc = lib().orc_pan_to_rle(_p(pan), H, W, _p(labels_a), labels_a.size, int(label_divisor),
                                  _p(things), things.size, int(bool(force_connected)),
                                  _p(inst), inst_cap, _p(runs), runs_cap,
                                  ctypes.byref(ni), ctypes.byref(nr))
        if rc == 0:
            break
        inst_cap *= 4
        runs_cap *= 4
    out = {int(l): {} for l in labels}
    r0 = 0
    for row in inst[:ni.value]:
        cls, lab, y0, x0, y1, x1, n = (int(v) for v in row)
        out[cls][lab] = {'box': (y0, x0, y1, x1),
                         'starts': runs[r0:r0 + n, 0].copy(),
                         'runs': runs[r0:r0 + n, 1].copy()}
        r0 += n
    return out


def rle_seg_to_pan_seg(rle_seg, shape):
    """rle.py:88-118."""
    pan = np.zeros(shape, np.uint32).ravel()
    for attrs in rle_seg.values():
        for oid, a in attrs.items():
            for s, r in zip(a['starts'], a['runs']):
                pan[s:s + r] = oid
    return pan.reshape(shape)
