"""ctypes binding of libempanada_b200.so (the C ABI declared in include/empanada_b200.h).

There is no CPU fallback: if the library is missing this module raises, loudly.  Build it with
``python -m empanada_b200.build`` (nvcc cross-compiles sm_100a without a GPU).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('EMP_B200_LIB') or os.path.join(_HERE, 'lib', 'libempanada_b200.so')

EMP_OK = 0
ST_K, ST_FLAGS, ST_NROWRUNS, ST_NRUNS, ST_NINST, ST_WORDS = 0, 1, 2, 3, 4, 16
FLAG_K_OVERFLOW, FLAG_CLASS_RANGE, FLAG_ID_RANGE, FLAG_RLE_OVERFLOW = 1, 2, 4, 8
MAX_THINGS, MAX_CLASSES, MAX_LABELS = 16, 4096, 64
# packed output of emp_stack_block (include/empanada_b200.h: EMP_BLK_*)
BLK_HDR_MAXLAB, BLK_HDR_WORDS, BLK_SLICE_WORDS, BLK_INST_WORDS = 4, 4 + 64, 6, 9
BLK_MAXLAB_OVERFLOW = 1 << 62           # emp_stack_blocks raises maxlab_all[0] to this if a slice overflowed a table

_lib = None

_vp, _i32, _i64, _f32, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t

_SIGNATURES = {
    'emp_version': (_i32, []),
    'emp_last_error': (ctypes.c_char_p, []),
    'emp_profile_enable': (_i32, [_i32]),
    'emp_profile_read': (_i32, [_vp, _vp]),
    'emp_workspace_bytes': (_sz, [_i32, _i32, _i32, _i32]),
    'emp_find_centers': (_i32, [_vp, _i32, _i32, _f32, _i32, _vp, _i32, _vp, _sz, _vp]),
    'emp_group_pixels': (_i32, [_vp, _i32, _vp, _i32, _i32, _f32, _i32, _vp, _i32, _vp, _sz, _vp]),
    'emp_instance_segmentation': (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _i32, _f32, _i32, _vp, _vp,
                                         _i32, _i32, _vp, _sz, _vp]),
    'emp_merge': (_i32, [_vp, _vp, _i32, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _sz, _vp]),
    'emp_merge_coarse': (_i32, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp, _i32, _i64,
                                _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    'emp_coarse_ids': (_i32, [_vp, _vp, _i32, _i32, _f32, _i32, _f32, _vp, _i32, _vp, _sz, _vp]),
    'emp_panoptic_batched': (_i32, [_i32, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _i32, _i64, _i64, _i64,
                                    _f32, _i32, _vp, _vp, _i32, _i32, _vp, _sz, _vp]),
    'emp_host_scratch_bytes': (_sz, [_i32, _i32, _i32, _i32]),
    'emp_host_sem_bytes_per_px': (ctypes.c_double, []),
    'emp_panoptic_batched_host': (_i32, [_i32, _vp, _vp, _vp, _i32, _i32, _vp, _i32, _i64, _i64, _i64,
                                         _f32, _i32, _vp, _vp, _vp, _i32, _vp, _sz]),
    'emp_panoptic_batched_host_u8': (_i32, [_i32, _vp, _vp, _vp, _i32, _i32, _vp, _i32, _i64, _i64, _i64,
                                            _f32, _i32, _vp, _vp, _vp, _i32, _vp, _sz]),
    'emp_median_harden': (_i32, [_vp, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _i32, _vp]),
    'emp_median_chain': (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _sz, _vp, _f32, _vp, _sz, _vp, _vp, _vp, _vp]),
    'emp_median_chain_repair': (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _sz, _vp, _vp, _f32, _vp, _sz, _vp, _vp, _vp, _vp]),
    'emp_rle_workspace_bytes': (_sz, [_i32, _i32, _i32, _i32, _i64]),
    'emp_rle': (_i32, [_vp, _i32, _i32, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _i32, _vp, _i32,
                       _vp, _sz, _vp]),
    'emp_stack_block_scratch_bytes': (_sz, [_vp, _i32]),
    'emp_stack_block_packed_words': (_sz, [_vp, _i32]),
    'emp_stack_block': (_i32, [_vp, _i32, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    'emp_stack_blocks': (_i32, [_vp, _i32, _i32, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _sz,
                                _vp, _vp, _vp]),
    'emp_take_slices': (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    'emp_rle_pair_overlaps': (_i32, [_vp, _sz, _vp, _i32, _i32, _vp, _i32, _vp, _vp]),
    'emp_rle_list_overlaps': (_i32, [_vp, _i32, _i64, _vp, _i32, _vp, _i32, _vp, _vp]),
    'emp_fill_runs': (_i32, [_vp, _sz, _vp, _i32, _i32, _vp, _sz, _vp, _i32, _sz, _vp]),
}

EXPORTS = tuple(_SIGNATURES)


class StackCfg(ctypes.Structure):
    """emp_stack_cfg (include/empanada_b200.h)."""
    _fields_ = [(n, ctypes.c_int32) for n in ('H', 'W', 'h', 'w', 'shift', 'nms_kernel', 'k_cap', 'crop_h', 'crop_w',
                                             'n_things', 'n_labels', 'force_connected', 'run_cap', 'inst_cap')] + \
               [('nms_threshold', ctypes.c_float), ('step', ctypes.c_float),
                ('label_divisor', ctypes.c_int64), ('stuff_area', ctypes.c_int64), ('void_label', ctypes.c_int64),
                ('thing_list', ctypes.c_void_p), ('labels', ctypes.c_void_p)]


def lib():
    """The loaded library.  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: the CUDA library is required (no CPU fallback). '
                f'Build it with `python -m empanada_b200.build`.')
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class NeedMap(ctypes.Structure):
    """emp_need_map (include/empanada_b200.h)."""
    _fields_ = [('map', ctypes.c_void_p), ('stride', ctypes.c_size_t), ('W', ctypes.c_int32), ('shift', ctypes.c_int32),
                ('wc', ctypes.c_int32), ('thing_bits', ctypes.c_uint64)]


STAGES = ('nms_peaks', 'emit_centers', 'assign', 'build_lut', 'apply_lut', 'median_harden', 'rle_mark', 'rle_runs',
          'bin_centers', 'median_chain', 'rle_block_keys', 'rle_block_mark', 'rle_block_emit', 'rle_block_runs',
          'rle_block_pack', 'memset')


def profile_enable(on):
    check(lib().emp_profile_enable(int(on)))


def profile_read():
    """{stage: (total ms, launches)} since the last read (waits for the recorded events)."""
    ms = (ctypes.c_double * len(STAGES))()
    n = (ctypes.c_int * len(STAGES))()
    check(lib().emp_profile_read(ms, n))
    return {s: (ms[i], n[i]) for i, s in enumerate(STAGES)}


def check(rc):
    if rc != EMP_OK:
        msg = lib().emp_last_error().decode(errors='replace')
        raise RuntimeError(f'libempanada_b200 error {rc}: {msg}')


def require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError('empanada_b200 runs on CUDA tensors only (there is no CPU fallback); '
                               f'got a tensor on {t.device}')
    dev = tensors[0].device
    for t in tensors[1:]:
        if t.device != dev:
            raise RuntimeError(f'tensors on different devices: {dev} vs {t.device}')
    return dev


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def i64_array(values):
    values = [int(v) for v in values]
    return (ctypes.c_int64 * max(1, len(values)))(*values), len(values)


_ws_cache = {}


def workspace(device, nbytes, tag=''):
    """A cached uint8 workspace tensor per (device, stream, tag) — grown, never shrunk.  Reuse is
    safe because every user enqueues on the current stream of `device`."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def read_status(ws, n_tiles=1, stride=None):
    """Read the int32 status block(s) at the start of the workspace (synchronises)."""
    if n_tiles == 1:
        return ws[:ST_WORDS * 4].view(torch.int32).cpu()
    idx = (torch.arange(n_tiles, device=ws.device) * stride)[:, None] + torch.arange(ST_WORDS * 4, device=ws.device)[None]
    return ws[idx.reshape(-1)].view(torch.int32).reshape(n_tiles, ST_WORDS).cpu()
