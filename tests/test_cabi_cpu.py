"""CPU: the C-ABI library loads without a GPU, exports every symbol include/empanada_b200.h
declares, and its host-only helpers (workspace sizing, argument validation) behave.  No compute
call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from empanada_b200 import build, _cabi
    build.build()
    return _cabi.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'empanada_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(emp_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from empanada_b200 import _cabi
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
    assert sorted(_cabi.EXPORTS) == names, 'ctypes table and header disagree'


def test_version_and_workspace_sizes(lib):
    assert lib.emp_version() >= 100
    small = lib.emp_workspace_bytes(256, 256, 1024, 1)
    big = lib.emp_workspace_bytes(4096, 4096, 32768, 1)
    assert 0 < small < big
    # 2 B/px code map below 61440 centers, 4 B/px above
    assert lib.emp_workspace_bytes(4096, 4096, 70000, 1) - big > 4096 * 4096
    assert big % 256 == 0
    assert lib.emp_workspace_bytes(0, 10, 10, 1) == 0
    assert lib.emp_rle_workspace_bytes(2048, 2048, 1 << 16, 3, 20000) > 0
    assert lib.emp_host_scratch_bytes(1024, 1024, 4096, 1) > 3 * 28 * 1024 * 1024


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected before any CUDA call, so this runs on a GPU-less box."""
    things = (ctypes.c_int64 * 20)(*range(20))
    rc = lib.emp_panoptic_batched(1, None, 0, None, None, 64, 64, things, 1, 1000, 0, 0, 0.1, 7, None, None, 0, 16,
                                  None, 0, None)
    assert rc == 1 and b'null' in lib.emp_last_error()
    dummy = ctypes.c_void_p(256)
    rc = lib.emp_merge(dummy, dummy, 64, 64, 1000, things, 20, 0, 0, 10, dummy, dummy, 1 << 30, None)
    assert rc == 1 and b'thing classes' in lib.emp_last_error()
    rc = lib.emp_find_centers(dummy, 64, 64, 0.1, 0, None, 0, dummy, 1 << 30, None)
    assert rc == 1
    rc = lib.emp_find_centers(dummy, 64, 64, 0.1, 7, None, 16, ctypes.c_void_p(256), 16, None)
    assert rc == 3          # workspace too small
    rc = lib.emp_median_harden(None, 4, 1, 8, 8, 0.5, None, None, 0, None)
    assert rc == 1
    labels = (ctypes.c_int64 * 2)(1, 1)
    rc = lib.emp_rle(dummy, 8, 8, labels, 2, 1000, things, 1, 1, dummy, 16, dummy, 16, dummy, 1 << 30, None)
    assert rc == 1 and b'duplicate' in lib.emp_last_error()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'empanada_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(import|from)\s+oracle\b', text, flags=re.M), f
                assert '/root/reference' not in text, f


def test_cpu_tensor_is_refused():
    import torch
    from empanada_b200.inference import postprocess as pp
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        pp.get_panoptic_segmentation(torch.zeros(1, 1, 8, 8, dtype=torch.int64), torch.zeros(1, 1, 8, 8),
                                     torch.zeros(1, 2, 8, 8), [1], 1000, 0, 0)


def test_rle_host_grouping_matches_oracle():
    """tables_to_rle_seg (host half of pan_seg_to_rle_seg) on tables built from the oracle's dict:
    runs handed over in ascending-start order with slot ids, exactly what emp_rle emits."""
    import oracle
    from empanada_b200.inference import rle
    from conftest import load_golden
    g = load_golden('rle_wrap_fc')
    p = g['params']
    want = oracle.pan_seg_to_rle_seg(g['in_pan'], p['labels'], p['label_divisor'], p['thing_list'], True)
    inst, runs = [], []
    for cls, attrs in want.items():
        for lab, a in attrs.items():
            slot = len(inst)
            inst.append([cls, lab, *a['box'], len(a['starts']), 0])
            runs += [[s, r, slot] for s, r in zip(a['starts'], a['runs'])]
    inst = np.asarray(inst, np.int64)
    runs = np.asarray(sorted(runs), np.int64)
    got = rle.tables_to_rle_seg(inst, runs, p['labels'])
    assert list(got.keys()) == list(want.keys())
    for cls in want:
        assert list(got[cls].keys()) == list(want[cls].keys())
        for lab in want[cls]:
            assert got[cls][lab]['box'] == want[cls][lab]['box']
            np.testing.assert_array_equal(got[cls][lab]['starts'], want[cls][lab]['starts'])
            np.testing.assert_array_equal(got[cls][lab]['runs'], want[cls][lab]['runs'])
