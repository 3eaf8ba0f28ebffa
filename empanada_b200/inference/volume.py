"""
Slices of a volume that lives in HBM: the GPU side of ``array_utils.take`` (reference array_utils.py:6-23) as
``VolumeDataset.__getitem__`` uses it (data/volume_dataset.py:37-53) to feed the stack and orthoplane loops of
scripts/pdl_inference3d.py:110-176.  ``take_slices`` returns the n slices i0 .. i0+n-1 along ``axis`` as one contiguous
(n, A, B) CUDA tensor; xy slices are one copy, xz slices row copies, and yz slices — whose elements lie a whole row
apart — are gathered n at a time through shared memory (libempanada_b200 ``emp_take_slices``).
"""
import ctypes

import torch

from empanada_b200 import _cabi as C

__all__ = ['take_slices']


def take_slices(volume, axis, i0, n, out=None):
    """volume: (D,H,W) contiguous CUDA tensor of uint8 / int8 or a 4-byte dtype; returns (n, A, B) of the same dtype."""
    dev = C.require_cuda(volume)
    if volume.dim() != 3 or not volume.is_contiguous():
        raise ValueError('take_slices expects a contiguous (D, H, W) volume')
    eb = volume.element_size()
    if eb not in (1, 4):
        raise TypeError(f'volume elements must be 1 or 4 bytes (got {volume.dtype})')
    D, H, W = (int(v) for v in volume.shape)
    if axis not in (0, 1, 2) or i0 < 0 or n < 1 or i0 + n > volume.shape[axis]:
        raise IndexError(f'slices {i0} .. {i0 + n} along axis {axis} of a volume of shape {tuple(volume.shape)}')
    shape = (n, H, W) if axis == 0 else (n, D, W) if axis == 1 else (n, D, H)
    if out is None:
        out = torch.empty(shape, dtype=volume.dtype, device=dev)
    assert out.is_contiguous() and tuple(out.shape) == shape and out.dtype == volume.dtype
    with torch.cuda.device(dev):
        C.check(C.lib().emp_take_slices(ctypes.c_void_p(volume.data_ptr()), eb, D, H, W, int(axis), int(i0), int(n),
                                        ctypes.c_void_p(out.data_ptr()), C.stream_ptr(dev)))
    return out
