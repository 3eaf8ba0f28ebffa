"""
Drop-in for ``empanada.inference.engines`` (reference empanada/inference/engines.py): same class
names, constructor keywords, methods and return shapes.  The CNN forward (``infer``) is untouched
PyTorch/cuDNN; everything after it — the recursive median queue, hardening, center finding, pixel
grouping, the semantic/instance merge and (for the Render engines) the nearest upsample of the
coarse instance cells — runs in libempanada_b200's sm_100a kernels.
"""
import ctypes
import math
from collections import deque

import torch
import torch.nn.functional as F

from empanada_b200 import _cabi as C
from empanada_b200.inference import postprocess as pp
from empanada_b200.inference.postprocess import (
    factor_pad, find_instance_center, group_pixels,
    get_instance_segmentation,
    merge_semantic_and_instance,
    get_panoptic_segmentation
)

__all__ = [
    'PanopticDeepLabEngine',
    'PanopticDeepLabEngine3d',
    'PanopticDeepLabRenderEngine',
    'PanopticDeepLabRenderEngine3d',
    'BCEngine', 'BCEngine3d',
]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


@torch.no_grad()
def logits_to_prob(logits):
    """softmax (multiclass) or sigmoid (binary) — engines.py:22-30; stays a torch op."""
    if logits.size(1) > 1:
        return F.softmax(logits, dim=1)
    return torch.sigmoid(logits)


def median_harden(planes, confidence_thr, want_median=True, want_sem=None):
    """libempanada_b200 ``emp_median_harden`` on a list of ks (odd) equal-shape (1,C,H,W) float32
    CUDA tensors.  Returns (median (1,C,H,W) or None, sem or None) where sem is (1,1,H,W) int64
    (``want_sem='i64'``) or uint8 (``'u8'``)."""
    ks = len(planes)
    dev = C.require_cuda(*planes)
    planes = [p.detach().to(torch.float32).contiguous() for p in planes]
    n, c, h, w = planes[0].shape
    assert n == 1
    median = torch.empty_like(planes[0]) if want_median else None
    sem = None
    if want_sem == 'i64':
        sem = torch.empty((1, 1, h, w), dtype=torch.int64, device=dev)
    elif want_sem == 'u8':
        sem = torch.empty((1, 1, h, w), dtype=torch.uint8, device=dev)
    ptrs = (ctypes.c_void_p * ks)(*[p.data_ptr() for p in planes])
    with torch.cuda.device(dev):
        C.check(C.lib().emp_median_harden(ptrs, ks, c, h, w, float(confidence_thr),
                                          _ptr(median) if median is not None else None,
                                          _ptr(sem) if sem is not None else None,
                                          int(want_sem == 'u8'), C.stream_ptr(dev)))
    return median, sem


class _Engine:
    def __init__(self, model):
        self.model = model.eval()

    def infer(self, image):
        raise NotImplementedError

    def to_model_device(self, tensor):
        device = next(self.model.parameters()).device
        return tensor.to(device, non_blocking=True)

    def __call__(self, image):
        raise NotImplementedError


class _MedianQueue:
    """A sliding window of the last ``median_kernel_size`` model outputs (reference engines.py:47-90).

    While the window is still shorter than half the kernel the newest entry passes through raw; until
    it is full nothing is emitted; once full, the middle entry is emitted with the requested tensors
    replaced by the per-element median over the window — and the replacement is stored back into the
    window, which is what makes the reference's filter recursive (later windows see filtered planes)."""

    def __init__(self, median_kernel_size, **kwargs):
        super().__init__(**kwargs)
        assert median_kernel_size % 2 == 1, "Kernel size must be odd integer!"
        self.ks = median_kernel_size
        self.mid_idx = median_kernel_size // 2
        self.reset()

    def reset(self):
        self.median_queue = deque(maxlen=self.ks)

    def enqueue(self, item):
        self.median_queue.append(item)

    @torch.no_grad()
    def get_median(self, key):
        window = [entry[key] for entry in self.median_queue]
        return median_harden(window, 0.0)[0]

    def get_next(self, keys):
        filled = len(self.median_queue)
        if filled <= self.mid_idx:
            return self.median_queue[-1]                # start of the stack: raw
        if filled < self.ks:
            return None                                 # still filling
        centre = self.median_queue[self.mid_idx]
        centre.update({key: self.get_median(key) for key in keys})
        return centre

    def end(self):
        return [self.median_queue[i] for i in range(self.mid_idx + 1, len(self.median_queue))]


def _single_image(image):
    assert image.ndim == 4 and image.size(0) == 1


def _log2_factor(upsampling):
    assert math.log(upsampling, 2).is_integer(), "Upsampling factor not log base 2!"
    return int(2 + math.log(upsampling, 2))


class PanopticDeepLabEngine(_Engine):
    def __init__(self, model, thing_list, label_divisor=1000, stuff_area=64, void_label=0,
                 nms_threshold=0.1, nms_kernel=7, confidence_thr=0.5, **kwargs):
        super().__init__(model=model)
        for name, value in (('thing_list', thing_list), ('label_divisor', label_divisor), ('stuff_area', stuff_area),
                            ('void_label', void_label), ('nms_threshold', nms_threshold), ('nms_kernel', nms_kernel),
                            ('confidence_thr', confidence_thr)):
            setattr(self, name, value)

    @torch.no_grad()
    def _harden_seg(self, sem):
        """(N,C,H,W) probabilities -> (N,1,H,W) int64: argmax over C, or >= confidence_thr."""
        outs = [median_harden([sem[i:i + 1]], self.confidence_thr, want_median=False, want_sem='i64')[1]
                for i in range(sem.size(0))]
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    @torch.no_grad()
    def infer(self, image):
        model_out = self.model(image)
        model_out['sem'] = logits_to_prob(model_out['sem_logits'])   # sem is NOT sem_logits
        return model_out

    @torch.no_grad()
    def postprocess(self, sem, ctr_hmp, offsets):
        pan_seg, _ = get_panoptic_segmentation(
            sem, ctr_hmp, offsets, self.thing_list,
            self.label_divisor, self.stuff_area,
            self.void_label, self.nms_threshold, self.nms_kernel
        )
        return pan_seg

    def _pan_from(self, heads):
        return self.postprocess(self._harden_seg(heads['sem']), heads['ctr_hmp'], heads['offsets'])

    def __call__(self, image):
        _single_image(image)
        return self._pan_from(self.infer(self.to_model_device(image)))


class PanopticDeepLabEngine3d(_MedianQueue, PanopticDeepLabEngine):
    def __init__(self, model, thing_list, label_divisor=1000, stuff_area=64, void_label=0,
                 nms_threshold=0.1, nms_kernel=7, confidence_thr=0.5, median_kernel_size=3, **kwargs):
        super().__init__(
            model=model, thing_list=thing_list, label_divisor=label_divisor,
            stuff_area=stuff_area, void_label=void_label,
            nms_threshold=nms_threshold, nms_kernel=nms_kernel,
            confidence_thr=confidence_thr, median_kernel_size=median_kernel_size,
            **kwargs
        )

    def end(self):
        """The slices still waiting behind the window's centre, post-processed raw (engines.py:183-198)."""
        return [self._pan_from(heads) for heads in _MedianQueue.end(self)]

    def __call__(self, image):
        _single_image(image)
        self.enqueue(self.infer(self.to_model_device(image)))
        heads = self.get_next(keys=['sem'])
        return None if heads is None else self._pan_from(heads)


class PanopticDeepLabRenderEngine(PanopticDeepLabEngine):
    def __init__(self, model, thing_list, label_divisor=1000, stuff_area=64, void_label=0,
                 nms_threshold=0.1, nms_kernel=7, confidence_thr=0.5, padding_factor=16,
                 coarse_boundaries=True, **kwargs):
        super().__init__(
            model=model, thing_list=thing_list,
            label_divisor=label_divisor, stuff_area=stuff_area, void_label=void_label,
            nms_threshold=nms_threshold, nms_kernel=nms_kernel,
            confidence_thr=confidence_thr
        )
        self.padding_factor = padding_factor
        self.coarse_boundaries = coarse_boundaries

    @torch.no_grad()
    def infer(self, image, render_steps=2):
        model_out = self.model(image, render_steps, interpolate_ins=not self.coarse_boundaries)
        model_out['sem'] = logits_to_prob(model_out['sem_logits'])
        return model_out

    # -- reference-shaped public methods (materialise the upsampled float cells) ---------------
    @torch.no_grad()
    def get_instance_cells(self, ctr_hmp, offsets, upsampling=1):
        """(1,1,h,w) heat-map + (1,2,h,w) offsets -> (1,1,h*s,w*s) float32 ids, s = upsampling*step
        (engines.py:257-275)."""
        step = 4 if self.coarse_boundaries else 1
        ids = self._coarse_ids(ctr_hmp, offsets, step)[0]            # (h,w) int32
        s = int(upsampling * step)
        cells = ids.to(torch.float32)[None, None]
        if s > 1:
            cells = cells.repeat_interleave(s, dim=2).repeat_interleave(s, dim=3)
        return cells

    @torch.no_grad()
    def get_panoptic_seg(self, sem, instance_cells):
        """sem (1,H,W) int64 + float cells (1,1,H,W) -> pan (1,H,W) int64 (engines.py:277-292)."""
        thing = torch.zeros_like(sem)
        for thing_class in self.thing_list:
            thing[sem == thing_class] = 1
        instance_seg = (thing * instance_cells[0]).long()
        return merge_semantic_and_instance(sem, instance_seg, self.label_divisor, self.thing_list,
                                           self.stuff_area, self.void_label)

    @torch.no_grad()
    def postprocess(self, sem, instance_cells):
        sem = self._harden_seg(sem)[0]
        return self.get_panoptic_seg(sem, instance_cells)

    # -- fused path used by __call__: coarse int32 ids, upsample folded into the merge ----------
    def _coarse_ids(self, ctr_hmp, offsets, step, k_cap=None):
        dev = C.require_cuda(ctr_hmp, offsets)
        hm = ctr_hmp.detach().to(torch.float32).contiguous()
        off = offsets.detach().to(torch.float32).contiguous()
        assert hm.size(0) == 1 and off.size(0) == 1
        h, w = hm.shape[-2:]
        L = C.lib()
        k_cap = min(k_cap or pp.DEFAULT_K_CAP, h * w)
        nbytes = L.emp_workspace_bytes(h, w, k_cap, 1)
        ws = C.workspace(dev, nbytes, 'coarse')
        ids = torch.empty((h, w), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            C.check(L.emp_coarse_ids(_ptr(hm), _ptr(off), h, w, float(self.nms_threshold), int(self.nms_kernel),
                                     float(step), _ptr(ids), k_cap, _ptr(ws), ws.numel(), C.stream_ptr(dev)))
        return ids, ws, k_cap

    @torch.no_grad()
    def _fused_enqueue(self, sem_prob, ctr_hmp, offsets, upsampling, k_cap=None, sem8=None):
        """Enqueue median-queue output -> pan (1,H,W) int64 on the current stream, with no dense
        intermediate besides uint8 sem and NO host synchronisation.  Returns (pan, coarse workspace,
        merge workspace, k_cap); the status blocks at the start of the two workspaces (K / overflow
        flag, class-range flags) are valid once the stream has run — and only until the next call
        reuses the cached workspaces, so callers that defer the check copy them out first."""
        step = 4 if self.coarse_boundaries else 1
        if sem8 is None:
            _, sem = median_harden([sem_prob], self.confidence_thr, want_median=False, want_sem='u8')
        else:
            sem = sem8                                  # already hardened (the z-block chain of inference/stack.py)
        dev = sem.device
        H, W = sem.shape[-2:]
        h, w = ctr_hmp.shape[-2:]
        s = int(upsampling * step)
        shift = int(math.log2(s))
        assert (1 << shift) == s
        L = C.lib()
        ids, cws, k_cap = self._coarse_ids(ctr_hmp, offsets, step, k_cap)
        things, nt = C.i64_array(self.thing_list)
        nbytes = L.emp_workspace_bytes(H, W, k_cap, max(nt, 1))
        ws = C.workspace(dev, nbytes, 'merge')
        pan = torch.empty((1, H, W), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            C.check(L.emp_merge_coarse(_ptr(sem), 1, _ptr(ids), h, w, shift, H, W, int(self.label_divisor),
                                       things, nt, int(self.stuff_area), int(self.void_label), k_cap,
                                       _ptr(cws), _ptr(pan), _ptr(ws), ws.numel(), C.stream_ptr(dev)))
        return pan, cws, ws, k_cap

    @torch.no_grad()
    def _fused_postprocess(self, sem_prob, ctr_hmp, offsets, upsampling, sem8=None):
        """median-queue output -> pan (1,H,W) int64; reads the status blocks back (one sync) and
        retries with a larger center table in the (rare) overflow case."""
        k_cap = None
        while True:
            pan, cws, ws, k_cap = self._fused_enqueue(sem_prob, ctr_hmp, offsets, upsampling, k_cap, sem8)
            st = C.read_status(cws)
            if not (int(st[C.ST_FLAGS]) & C.FLAG_K_OVERFLOW):
                break
            k_cap = int(st[C.ST_K])
        pp._check_flags(int(C.read_status(ws)[C.ST_FLAGS]))
        return pan

    def _heads(self, image, upsampling):
        """Checks, pads to the model's stride, runs the CNN with the render steps the upsampling needs."""
        steps = _log2_factor(upsampling)
        _single_image(image)
        return self.infer(self.to_model_device(factor_pad(image, self.padding_factor)), steps)

    def _pan_cropped(self, heads, size, upsampling):
        pan = self._fused_postprocess(heads['sem'], heads['ctr_hmp'], heads['offsets'], upsampling)
        return pan[..., :size[0], :size[1]]

    def __call__(self, image, size, upsampling=1):
        return self._pan_cropped(self._heads(image, upsampling), size, upsampling)


class PanopticDeepLabRenderEngine3d(_MedianQueue, PanopticDeepLabRenderEngine):
    def __init__(self, model, thing_list, label_divisor=1000, stuff_area=64, void_label=0,
                 nms_threshold=0.1, nms_kernel=7, confidence_thr=0.5, median_kernel_size=3,
                 padding_factor=16, coarse_boundaries=True, **kwargs):
        super().__init__(
            model=model, thing_list=thing_list,
            label_divisor=label_divisor, stuff_area=stuff_area, void_label=void_label,
            nms_threshold=nms_threshold, nms_kernel=nms_kernel,
            confidence_thr=confidence_thr, median_kernel_size=median_kernel_size,
            padding_factor=padding_factor, coarse_boundaries=coarse_boundaries
        )

    def end(self, upsampling=1):
        return [self._pan_cropped(heads, heads['size'], upsampling) for heads in _MedianQueue.end(self)]

    def __call__(self, image, size, upsampling=1):
        heads = self._heads(image, upsampling)
        heads['size'] = size
        self.enqueue(heads)
        heads = self.get_next(keys=['sem'])
        return None if heads is None else self._pan_cropped(heads, size, upsampling)


def _boundary_contour(model_out):
    """sigmoid of the (binary) semantic and contour logits, stacked: (N, 2, H, W)."""
    sem_logits, cnt_logits = model_out['sem_logits'], model_out['cnt_logits']
    assert sem_logits.size(1) == 1
    return {'bc': torch.sigmoid(torch.cat([sem_logits, cnt_logits], dim=1))}


class BCEngine(_Engine):
    """Boundary-contour models: sigmoid + concat only (engines.py:396-416); no panoptic post-proc."""

    def __init__(self, model, **kwargs):
        super().__init__(model=model)

    @torch.no_grad()
    def infer(self, image):
        return _boundary_contour(self.model(image))

    def __call__(self, image):
        _single_image(image)
        return self.infer(self.to_model_device(image))['bc']


class BCEngine3d(_MedianQueue, BCEngine):
    def __init__(self, model, median_kernel_size=3, padding_factor=16, **kwargs):
        super().__init__(model=model, median_kernel_size=median_kernel_size)
        self.padding_factor = padding_factor

    @torch.no_grad()
    def infer(self, image, render_steps=2):
        return _boundary_contour(self.model(image, render_steps))

    @staticmethod
    def _cropped(out):
        return out['bc'][..., :out['size'][0], :out['size'][1]]

    def end(self, upsampling=1):
        return [self._cropped(out) for out in _MedianQueue.end(self)]

    def __call__(self, image, size, upsampling=1):
        steps = _log2_factor(upsampling)
        _single_image(image)
        out = self.infer(self.to_model_device(factor_pad(image, self.padding_factor)), steps)
        out['size'] = size
        self.enqueue(out)
        out = self.get_next(keys=['bc'])
        return None if out is None else self._cropped(out)
