"""B200: median/harden and pan_seg -> RLE kernels against reference-generated fixtures and the
oracle."""
import numpy as np
import pytest
import torch

import oracle
from conftest import golden_names, load_golden
from empanada_b200.inference import engines as eng
from empanada_b200.inference import rle
from empanada_b200.synth import synth_tile

pytestmark = pytest.mark.gpu


def cu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize('ks', [1, 3, 5, 7, 9, 11])
@pytest.mark.parametrize('shape', [(1, 1, 64, 96), (1, 3, 33, 47), (1, 2, 128, 130)])
def test_median_harden_vs_oracle(ks, shape, cuda_device):
    rng = np.random.default_rng(ks * 100 + shape[1])
    planes = [rng.random(shape, dtype=np.float32) for _ in range(ks)]
    for p in planes[1:]:
        m = rng.random(shape) < 0.2            # force ties between planes
        p[m] = planes[0][m]
    want_med = oracle.median_planes(planes)
    want_sem = oracle.harden_seg(want_med, 0.3)
    for fmt, dt in (('i64', torch.int64), ('u8', torch.uint8)):
        med, sem = eng.median_harden([cu(p, cuda_device) for p in planes], 0.3, True, fmt)
        np.testing.assert_array_equal(med.cpu().numpy(), want_med)
        assert sem.dtype == dt and tuple(sem.shape) == want_sem.shape
        np.testing.assert_array_equal(sem.cpu().numpy().astype(np.int64), want_sem)
    # torch's own median agrees too (middle order statistic)
    tmed = torch.median(torch.cat([torch.from_numpy(p) for p in planes], 0), dim=0, keepdim=True).values
    np.testing.assert_array_equal(tmed.numpy(), want_med)


def test_harden_threshold_edges(cuda_device):
    thr = 0.3
    t32 = np.float32(thr)
    vals = np.array([np.nextafter(t32, np.float32(0)), t32, np.nextafter(t32, np.float32(1)), 0.0, 1.0], np.float32)
    prob = np.tile(vals, 8).reshape(1, 1, 5, 8)
    _, sem = eng.median_harden([cu(prob, cuda_device)], thr, False, 'i64')
    np.testing.assert_array_equal(sem.cpu().numpy(), oracle.harden_seg(prob, thr))
    np.testing.assert_array_equal(sem.cpu().numpy(), (torch.from_numpy(prob) >= thr).long().numpy())


def _flatten(seg):
    inst, starts, runs = [], [], []
    for cls, attrs in seg.items():
        for lab, a in attrs.items():
            assert isinstance(a['box'], tuple) and a['starts'].dtype == np.int64 and a['runs'].dtype == np.int64
            inst.append([cls, lab, *a['box'], len(a['starts'])])
            starts += list(a['starts'])
            runs += list(a['runs'])
    return (np.asarray(inst, np.int64).reshape(-1, 7), np.asarray(starts, np.int64), np.asarray(runs, np.int64))


@pytest.mark.parametrize('name', golden_names('rle_'))
def test_golden_rle(name, cuda_device):
    g = load_golden(name)
    p = g['params']
    for src in (g['in_pan'], cu(g['in_pan'], cuda_device)):       # numpy (reference style) and CUDA tensor
        seg = rle.pan_seg_to_rle_seg(src, p['labels'], p['label_divisor'], p['thing_list'], p['force_connected'])
        assert list(seg.keys()) == p['labels']
        inst, starts, runs = _flatten(seg)
        np.testing.assert_array_equal(inst, g['out_inst'])
        np.testing.assert_array_equal(starts, g['out_starts'])
        np.testing.assert_array_equal(runs, g['out_runs'])
        np.testing.assert_array_equal(rle.rle_seg_to_pan_seg(seg, g['in_pan'].shape), g['out_back'])


@pytest.mark.parametrize('fc', [True, False])
@pytest.mark.parametrize('case', [(257, 300, 60, 1), (1024, 1024, 500, 2), (96, 2048, 200, 3)])
def test_rle_vs_oracle(case, fc, cuda_device):
    H, W, n, seed = case
    d = synth_tile(H, W, n, seed, semi_axes=(4, 16), sigma=3.0, thing_classes=(1, 2), stuff_classes=(3,))
    pan, _ = oracle.get_panoptic_segmentation(d['sem'], d['ctr_hmp'], d['offsets'], [1, 2], 1000, 16, 0, 0.1, 5)
    pan = pan[0, 0]
    want = oracle.pan_seg_to_rle_seg(pan, [1, 2, 3], 1000, [1, 2], fc)
    got = rle.pan_seg_to_rle_seg(cu(pan, cuda_device), [1, 2, 3], 1000, [1, 2], fc)
    wi, ws, wr = _flatten(want)
    gi, gs, gr = _flatten(got)
    np.testing.assert_array_equal(gi, wi)
    np.testing.assert_array_equal(gs, ws)
    np.testing.assert_array_equal(gr, wr)
    if not fc:          # without CCL relabelling the codec is lossless
        np.testing.assert_array_equal(rle.rle_seg_to_pan_seg(got, pan.shape), pan.astype(np.uint32))


def test_rle_capacity_retry_and_noise(cuda_device):
    """Salt-and-pepper map: more row-runs than the initial capacity -> overflow flag -> retry."""
    rng = np.random.default_rng(5)
    pan = np.where(rng.random((512, 640)) < 0.5, 1001, 0).astype(np.int64)
    pan[rng.random(pan.shape) < 0.1] = 2000
    want = oracle.pan_seg_to_rle_seg(pan, [1, 2], 1000, [1], True)
    inst, runs = rle.rle_tables(cu(pan, cuda_device), [1, 2], 1000, [1], True, run_cap=4096)
    got = rle.tables_to_rle_seg(inst, runs, [1, 2])
    for a, b in zip(_flatten(got), _flatten(want)):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize('ks,C', [(3, 1), (5, 1), (3, 3)])
def test_median_harden_propagates_nan(ks, C, cuda_device):
    """torch.median gives NaN for a window holding one (engines.py:59-66 runs it on the queued planes), argmax treats
    NaN as the maximum, `>= thr` is false for it: the kernel must agree plane for plane."""
    rng = np.random.default_rng(ks * 10 + C)
    planes = [rng.random((1, C, 17, 23), dtype=np.float32) for _ in range(ks)]
    planes[1][0, 0, 3, 4] = np.nan
    planes[ks - 1][0, C - 1, 9, 9] = np.nan
    planes[0][0, 0, 9, 9] = np.nan
    t = [torch.from_numpy(p).to(cuda_device) for p in planes]
    med, sem = eng.median_harden(t, 0.5, want_median=True, want_sem='i64')
    want = torch.median(torch.cat(t, dim=0), dim=0, keepdim=True).values
    assert torch.equal(torch.isnan(med), torch.isnan(want))
    assert torch.equal(torch.nan_to_num(med, nan=-1.0), torch.nan_to_num(want, nan=-1.0))
    want_sem = torch.argmax(want, dim=1, keepdim=True) if C > 1 else (want >= 0.5).long()
    assert torch.equal(sem, want_sem)
