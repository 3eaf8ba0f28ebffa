"""B200: the CUDA path, called through the reference-shaped Python API (which goes through the
C ABI), against (i) fixtures produced by the reference itself and (ii) the oracle on seeded
synthetic tiles.  Everything is bit-exact (integer labels / indices)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import golden_names, load_golden
from empanada_b200.inference import postprocess as pp
from empanada_b200.synth import synth_tile

pytestmark = pytest.mark.gpu


def cu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize('name', golden_names('pp_'))
def test_golden_find_centers(name, cuda_device):
    g = load_golden(name)
    p = g['params']
    ctr = pp.find_instance_center(cu(g['in_hm'], cuda_device), p['threshold'], p['nms_kernel'])
    assert ctr.dtype == torch.int64 and ctr.dim() == 2 and ctr.shape[1] == 2
    np.testing.assert_array_equal(ctr.cpu().numpy(), g['out_centers'])


@pytest.mark.parametrize('name', golden_names('pp_'))
def test_golden_group_pixels(name, cuda_device):
    g = load_golden(name)
    if g['out_centers'].shape[0] == 0:
        with pytest.raises(AssertionError):
            pp.group_pixels(cu(g['out_centers'], cuda_device), cu(g['in_off'], cuda_device))
        return
    for step in (1, 4):
        key = f'out_ids_step{step}'
        if key not in g:
            continue
        ids = pp.group_pixels(cu(g['out_centers'], cuda_device), cu(g['in_off'], cuda_device), step=float(step))
        assert ids.dtype == torch.int64 and tuple(ids.shape) == g[key].shape
        got = ids.cpu().numpy()
        bad = np.argwhere(got != g[key])
        assert bad.shape[0] == 0, f'{bad.shape[0]} mismatches, first {bad[:5]}, got {got[tuple(bad[0])]} want {g[key][tuple(bad[0])]}'


@pytest.mark.parametrize('name', golden_names('pp_'))
def test_golden_instance_and_panoptic(name, cuda_device):
    g = load_golden(name)
    p = g['params']
    sem, hm, off = (cu(g[k], cuda_device) for k in ('in_sem', 'in_hm', 'in_off'))
    ins, ctr = pp.get_instance_segmentation(sem, hm, off, p['thing_list'], p['threshold'], p['nms_kernel'])
    assert tuple(ins.shape) == g['out_ins'].shape and ctr.shape[0] == 1
    np.testing.assert_array_equal(ctr[0].cpu().numpy(), g['out_centers'])
    np.testing.assert_array_equal(ins.cpu().numpy(), g['out_ins'])
    pan, ctr = pp.get_panoptic_segmentation(sem, hm, off, p['thing_list'], p['label_divisor'], p['stuff_area'],
                                            p['void_label'], p['threshold'], p['nms_kernel'])
    assert pan.dtype == torch.int64 and tuple(pan.shape) == g['out_pan'].shape
    np.testing.assert_array_equal(ctr[0].cpu().numpy(), g['out_centers'])
    got = pan.cpu().numpy()
    bad = np.argwhere(got != g['out_pan'])
    assert bad.shape[0] == 0, f'{bad.shape[0]} mismatches, first {bad[:5]}'
    # the separate merge entry point on the reference's own instance map
    pan2 = pp.merge_semantic_and_instance(sem, cu(g['out_ins'], cuda_device), p['label_divisor'], p['thing_list'],
                                          p['stuff_area'], p['void_label'])
    assert tuple(pan2.shape) == g['out_pan'].shape
    np.testing.assert_array_equal(pan2.cpu().numpy(), g['out_pan'])
    # uint8 semantic input (the engines' internal format) gives the same map
    if g['in_sem'].min() >= 0 and g['in_sem'].max() < 256:
        H, W = g['in_sem'].shape[-2:]
        pan3, _, _ = pp._panoptic_tiles(sem.to(torch.uint8).reshape(1, H, W), hm.reshape(1, H, W).contiguous(),
                                        off.reshape(1, 2, H, W).contiguous(), p['thing_list'], p['label_divisor'],
                                        p['stuff_area'], p['void_label'], p['threshold'], p['nms_kernel'])
        np.testing.assert_array_equal(pan3.cpu().numpy().reshape(g['out_pan'].shape), g['out_pan'])


@pytest.mark.parametrize('name', golden_names('merge_'))
def test_golden_merge(name, cuda_device):
    g = load_golden(name)
    p = g['params']
    pan = pp.merge_semantic_and_instance(cu(g['in_sem'], cuda_device), cu(g['in_ins'], cuda_device),
                                         p['label_divisor'], p['thing_list'], p['stuff_area'], p['void_label'])
    assert tuple(pan.shape) == g['out_pan'].shape and pan.dtype == torch.int64
    np.testing.assert_array_equal(pan.cpu().numpy(), g['out_pan'])


CASES = [
    # H, W, n_inst, seed, semi_axes, sigma, things, stuff, nms_k, L, stuff_area, void
    (300, 417, 25, 100, (6, 20), 4.0, (1,), (), 7, 1000, 64, 0),          # odd width -> scalar path
    (512, 768, 120, 101, (6, 22), 4.0, (1,), (), 7, 1000, 64, 0),
    (1024, 1024, 500, 102, (5, 18), 3.0, (1,), (), 7, 1000, 64, 0),       # K ~ 500, many tiles
    (640, 512, 200, 103, (4, 14), 2.5, (1, 3), (2, 4), 5, 20000, 32, -1), # multi-class
    (256, 4096, 300, 104, (4, 12), 2.5, (2,), (1,), 3, 1000, 0, 7),       # wide
    (1000, 36, 30, 105, (3, 9), 2.0, (1,), (), 3, 1000, 16, 0),           # narrow
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: f'{c[0]}x{c[1]}_n{c[2]}')
def test_synthetic_vs_oracle(case, cuda_device):
    H, W, n, seed, axes, sigma, things, stuff, k, L, sa, void = case
    d = synth_tile(H, W, n, seed, semi_axes=axes, sigma=sigma, thing_classes=things, stuff_classes=stuff)
    want_pan, want_ctr = oracle.get_panoptic_segmentation(d['sem'], d['ctr_hmp'], d['offsets'], list(things),
                                                          L, sa, void, 0.1, k)
    sem, hm, off = (cu(d[k2], cuda_device) for k2 in ('sem', 'ctr_hmp', 'offsets'))
    pan, ctr = pp.get_panoptic_segmentation(sem, hm, off, list(things), L, sa, void, 0.1, k)
    np.testing.assert_array_equal(ctr[0].cpu().numpy(), want_ctr[0])
    got = pan.cpu().numpy()
    bad = np.argwhere(got != want_pan)
    assert bad.shape[0] == 0, f'{bad.shape[0]} mismatches of {got.size}, first {bad[:5]}'
    # standalone group_pixels over ALL pixels (background offsets are 0 -> far from any center)
    if want_ctr.shape[1] > 0 and H * W <= 1 << 20:
        want_ids = oracle.group_pixels(want_ctr[0], d['offsets'])
        ids = pp.group_pixels(ctr[0], off)
        np.testing.assert_array_equal(ids.cpu().numpy(), want_ids)


def test_random_offsets_defeat_the_cull(cuda_device):
    """Uniformly random offsets: every tile's location box covers the image, so nothing can be
    culled and the kernel degenerates to the exact brute-force argmin."""
    rng = np.random.default_rng(7)
    H, W = 96, 160
    hm = np.zeros((H, W), np.float32)
    ys, xs = rng.integers(0, H, 60), rng.integers(0, W, 60)
    hm[ys, xs] = rng.uniform(0.5, 1.0, 60).astype(np.float32)
    off = rng.uniform(-150, 150, (1, 2, H, W)).astype(np.float32)
    ctr = oracle.find_instance_center(hm, 0.1, 3)
    want = oracle.group_pixels(ctr, off)
    got = pp.group_pixels(cu(ctr, cuda_device), cu(off, cuda_device))
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_many_centers_chunked_candidates(cuda_device):
    """K far above the candidate-list capacity (1024) with nothing cullable."""
    rng = np.random.default_rng(8)
    H, W = 64, 96
    hm = (rng.random((H, W)) < 0.45).astype(np.float32) * rng.uniform(0.2, 1, (H, W)).astype(np.float32)
    ctr = oracle.find_instance_center(hm, 0.1, 1)
    assert ctr.shape[0] > 2048
    off = rng.uniform(-80, 80, (1, 2, H, W)).astype(np.float32)
    want = oracle.group_pixels(ctr, off)
    got_ctr = pp.find_instance_center(cu(hm[None, None], cuda_device), 0.1, 1)
    np.testing.assert_array_equal(got_ctr.cpu().numpy(), ctr)
    got = pp.group_pixels(got_ctr, cu(off, cuda_device))
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_nonfinite_offsets(cuda_device):
    rng = np.random.default_rng(9)
    H, W = 40, 64
    for K in (5, 30):
        hm = np.zeros((H, W), np.float32)
        hm[rng.integers(0, H, K), rng.integers(0, W, K)] = 1.0
        ctr = oracle.find_instance_center(hm, 0.1, 1)
        off = rng.normal(0, 3, (1, 2, H, W)).astype(np.float32)
        off[0, 0, 3, 5] = np.inf
        off[0, 1, 10, 7] = -np.inf
        off[0, 0, 20, 9] = np.nan
        off[0, 1, 33, 60] = 3e38
        want = oracle.group_pixels(ctr, off)
        got = pp.group_pixels(cu(ctr, cuda_device), cu(off, cuda_device))
        np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_error_conventions(cuda_device):
    dev = cuda_device
    sem = torch.zeros((1, 1, 8, 8), dtype=torch.int64, device=dev)
    hm = torch.zeros((1, 1, 8, 8), device=dev)
    off = torch.zeros((1, 2, 8, 8), device=dev)
    with pytest.raises(ValueError, match='single channel'):
        pp.get_panoptic_segmentation(sem.repeat(1, 2, 1, 1), hm, off, [1], 1000, 0, 0)
    with pytest.raises(ValueError, match='batch size = 1'):
        pp.get_panoptic_segmentation(sem.repeat(2, 1, 1, 1), hm, off, [1], 1000, 0, 0)
    with pytest.raises(ValueError, match='batch size = 1'):
        pp.get_panoptic_segmentation(sem, hm, off.repeat(2, 1, 1, 1), [1], 1000, 0, 0)
    with pytest.raises(ValueError, match='batch size = 1'):
        pp.group_pixels(torch.zeros((1, 2), dtype=torch.int64, device=dev), off.repeat(2, 1, 1, 1))
    with pytest.raises(AssertionError):
        pp.group_pixels(torch.zeros((0, 2), dtype=torch.int64, device=dev), off)
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        pp.find_instance_center(hm.cpu())
    # K == 0 is not an error in the fused path: everything is stuff / void
    pan, ctr = pp.get_panoptic_segmentation(sem, hm, off, [1], 1000, 0, 0)
    assert tuple(pan.shape) == (1, 1, 8, 8) and tuple(ctr.shape) == (1, 0, 2) and int(pan.abs().sum()) == 0
    # inputs are never modified
    hm2 = torch.rand((1, 1, 32, 32), device=dev)
    keep = hm2.clone()
    pp.find_instance_center(hm2, 0.5, 3)
    assert torch.equal(hm2, keep)
    assert pp.factor_pad(torch.zeros(1, 1, 30, 33), 16).shape == (1, 1, 32, 48)


def test_full_size_tile_matches_oracle_and_batching(cuda_device):
    """BASELINE config-2 shape: one 4096x4096 tile with ~500 centers against the oracle, then the
    batched entry point must give the same map for every copy in a batch."""
    H = W = 4096
    d = synth_tile(H, W, 500, seed=0)
    want_pan, want_ctr = oracle.get_panoptic_segmentation(d['sem'], d['ctr_hmp'], d['offsets'], [1], 1000, 64, 0, 0.1, 7)
    sem, hm, off = (cu(d[k], cuda_device) for k in ('sem', 'ctr_hmp', 'offsets'))
    pan, ctr = pp.get_panoptic_segmentation(sem, hm, off, [1], 1000, 64, 0, 0.1, 7)
    np.testing.assert_array_equal(ctr[0].cpu().numpy(), want_ctr[0])
    assert 350 <= ctr.shape[1] <= 520
    want = torch.from_numpy(want_pan).to(cuda_device)
    assert torch.equal(pan, want)
    B = 3
    pan_b, ctr_b, Ks = pp._panoptic_tiles(sem.reshape(1, H, W).repeat(B, 1, 1), hm.reshape(1, H, W).repeat(B, 1, 1),
                                          off.reshape(1, 2, H, W).repeat(B, 1, 1, 1), [1], 1000, 64, 0, 0.1, 7)
    assert Ks == [ctr.shape[1]] * B
    for b in range(B):
        assert torch.equal(pan_b[b], want[0, 0])
    # size-independent properties: labels of class 1 are 1001..1000+n with no gaps; stuff is 0
    labs = torch.unique(pan)
    things = labs[labs > 0]
    assert int(things.min()) == 1001 and int(things.max()) == 1000 + things.numel()
