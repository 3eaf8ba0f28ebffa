"""B200: cases aimed at the seams of the kernels' own structure — item / strip / cell borders, the TMA
staging conditions, batched launches with different tiles — all bit-exact against the oracle."""
import numpy as np
import pytest
import torch

import oracle
from empanada_b200.inference import postprocess as pp
from empanada_b200.synth import synth_tile

pytestmark = pytest.mark.gpu


def cu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize('W', [600, 601, 256, 255])
@pytest.mark.parametrize('k', [1, 2, 3, 4, 7])
def test_nms_plateaus_across_item_borders(W, k, cuda_device):
    """Equal-valued plateaus and near-peaks straddling the nms kernel's item borders (rows 3|4, 63|64,
    columns 255|256, 31|32), the image border and the TMA halo rows; W % 4 != 0 takes the unstaged path."""
    rng = np.random.default_rng(1000 + W + k)
    H = 140
    hm = rng.uniform(0.0, 0.09, (H, W)).astype(np.float32)          # below threshold noise
    for (y, x) in [(3, 255), (4, 256), (63, 31), (64, 32), (0, 0), (H - 1, W - 1), (67, 254), (3, 100), (4, 100),
                   (127, 200), (128, 201), (5, W - 1), (H - 1, 17)]:
        if x < W:
            hm[max(0, y - 1):y + 2, max(0, x - 1):x + 2] = 0.5                     # 3x3 plateau of ties
    ys, xs = rng.integers(0, H, 80), rng.integers(0, W, 80)
    hm[ys, xs] = rng.choice(np.array([0.3, 0.5, 0.7, 0.9], np.float32), 80)        # isolated peaks, many ties in value
    hm[10, 40:60] = 0.8                                                             # a ridge
    hm[rng.integers(0, H, 20), rng.integers(0, W, 20)] = -0.4                       # negatives never count
    for thr in (0.1, -0.2):
        want = oracle.find_instance_center(hm, thr, k)
        got = pp.find_instance_center(cu(hm[None, None], cuda_device), thr, k)
        np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize('step', [1.0, 4.0])
def test_locations_on_cell_borders_and_exact_ties(step, cuda_device):
    """Centers on a lattice and integer offsets: shifted locations sit exactly on cell borders of the
    center index and exactly between 2 or 4 centers, so the sqrt-free pass must hand the ties to the
    precise path (lowest index wins) and ring 0 must never settle a pixel wrongly."""
    H, W = 256, 320
    yy, xx = np.meshgrid(np.arange(8, H, 16), np.arange(8, W, 16), indexing='ij')
    ctr = np.stack([yy.ravel(), xx.ravel()], 1).astype(np.int64)                   # K = 320 > 20: chunked semantics
    rng = np.random.default_rng(5)
    off = np.zeros((1, 2, H, W), np.float32)
    off[0, 0] = rng.integers(-24, 25, (H, W)) * step                               # integer multiples of step: exact ties
    off[0, 1] = rng.integers(-24, 25, (H, W)) * step
    off[0, :, 100:120, 50:90] = 0.5 * step                                          # half-integer
    off[0, 0, 200:210, :] = 2e5                                                     # sentinel: farther than 1e5 -> id 0
    want = oracle.group_pixels(ctr, off, 20, step)
    got = pp.group_pixels(cu(ctr, cuda_device), cu(off, cuda_device), 20, step)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    assert (want == 0).any() and (want > 0).any()


def test_few_centers_no_sentinel(cuda_device):
    """K <= chunksize: no sentinel; one center only -> every pixel gets id 1, however far."""
    H, W = 64, 128
    off = np.random.default_rng(3).normal(0, 40, (1, 2, H, W)).astype(np.float32)
    off[0, 0, :8] = 3e5
    for ctr in (np.array([[10, 100]], np.int64), np.array([[0, 0], [63, 127], [30, 60]], np.int64)):
        want = oracle.group_pixels(ctr, off)
        got = pp.group_pixels(cu(ctr, cuda_device), cu(off, cuda_device))
        np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_batch_of_distinct_tiles(cuda_device):
    """One batched launch over different tiles: per-tile state (K, cell index, votes, areas, flags) must
    not leak between tiles when a persistent warp moves from one tile to the next."""
    H, W, B = 320, 448, 5
    tiles = [synth_tile(H, W, n, seed=300 + i, semi_axes=(5, 18), sigma=3.0, thing_classes=(1, 2), stuff_classes=(3,))
             for i, n in enumerate([40, 1, 90, 3, 60])]
    tiles[1]['sem'][:] = 3                                                          # a tile without any thing or center
    tiles[1]['ctr_hmp'][:] = 0
    sem = cu(np.stack([t['sem'][0, 0] for t in tiles]), cuda_device)
    hm = cu(np.stack([t['ctr_hmp'][0, 0] for t in tiles]), cuda_device)
    off = cu(np.stack([t['offsets'][0] for t in tiles]), cuda_device)
    pan, ctr, Ks = pp._panoptic_tiles(sem, hm, off, [1, 2], 1000, 64, -1, 0.1, 7)
    for b, t in enumerate(tiles):
        want_pan, want_ctr = oracle.get_panoptic_segmentation(t['sem'], t['ctr_hmp'], t['offsets'], [1, 2], 1000, 64, -1, 0.1, 7)
        assert Ks[b] == want_ctr.shape[1]
        np.testing.assert_array_equal(ctr[b, :Ks[b]].cpu().numpy(), want_ctr[0])
        np.testing.assert_array_equal(pan[b].cpu().numpy(), want_pan[0, 0])


@pytest.mark.parametrize('shape', [(192, 256), (200, 264), (64, 48), (3, 500)])
def test_uint8_sem_equals_int64_sem(shape, cuda_device):
    """sem as uint8 (our own _harden_seg output): TMA-staged when W % 16 == 0, scalar otherwise."""
    H, W = shape
    d = synth_tile(H, W, 30, seed=77, semi_axes=(4, 14), sigma=2.5, thing_classes=(1,), stuff_classes=(2,))
    want_pan, want_ctr = oracle.get_panoptic_segmentation(d['sem'], d['ctr_hmp'], d['offsets'], [1], 1000, 16, 0, 0.1, 5)
    hm, off = cu(d['ctr_hmp'], cuda_device), cu(d['offsets'], cuda_device)
    for dt in (np.int64, np.uint8):
        pan, ctr = pp.get_panoptic_segmentation(cu(d['sem'].astype(dt), cuda_device), hm, off, [1], 1000, 16, 0, 0.1, 5)
        np.testing.assert_array_equal(pan.cpu().numpy(), want_pan)
        np.testing.assert_array_equal(ctr.cpu().numpy(), want_ctr)


def test_unaligned_views_take_the_scalar_path(cuda_device):
    """Tensors whose storage is not 16-byte aligned (a sliced view made contiguous at an odd offset)."""
    H, W = 130, 260
    d = synth_tile(H, W, 25, seed=91, semi_axes=(4, 12), sigma=2.5)
    want_pan, want_ctr = oracle.get_panoptic_segmentation(d['sem'], d['ctr_hmp'], d['offsets'], [1], 1000, 64, 0, 0.1, 7)

    def odd(a):                     # same values, data_ptr shifted by one element
        t = torch.empty(a.size + 1, dtype=torch.from_numpy(a).dtype, device=cuda_device)
        v = t[1:].view(a.shape)
        v.copy_(torch.from_numpy(a))
        return v
    pan, ctr = pp.get_panoptic_segmentation(odd(d['sem']), odd(d['ctr_hmp']), odd(d['offsets']), [1], 1000, 64, 0, 0.1, 7)
    np.testing.assert_array_equal(pan.cpu().numpy(), want_pan)
    np.testing.assert_array_equal(ctr.cpu().numpy(), want_ctr)


def test_dense_small_instances(cuda_device):
    """BASELINE config 5 density (5000 instances of semi-axes 4..12 per 4096^2) on a 1024^2 crop: the
    center index runs with small cells, many strips hold several instances."""
    H = W = 1024
    d = synth_tile(H, W, 5000 // 16, seed=500, semi_axes=(4, 12), sigma=2.0)
    want_pan, want_ctr = oracle.get_panoptic_segmentation(d['sem'], d['ctr_hmp'], d['offsets'], [1], 1000, 64, 0, 0.1, 7)
    pan, ctr = pp.get_panoptic_segmentation(*(cu(d[k], cuda_device) for k in ('sem', 'ctr_hmp', 'offsets')), [1], 1000, 64, 0, 0.1, 7)
    np.testing.assert_array_equal(ctr.cpu().numpy(), want_ctr)
    assert ctr.shape[1] > 200
    np.testing.assert_array_equal(pan.cpu().numpy(), want_pan)


def test_dense_full_size_properties(cuda_device):
    """Config 5 at full size (4096^2, ~5000 instances) through size-independent properties: instance
    labels are 1001..1000+n without gaps, every thing pixel with a center within reach is labelled, the
    map is idempotent under a second run, and merge(sem, instance ids) reproduces the fused result."""
    H = W = 4096
    d = synth_tile(H, W, 5000, seed=501, semi_axes=(4, 12), sigma=2.0)
    sem, hm, off = (cu(d[k], cuda_device) for k in ('sem', 'ctr_hmp', 'offsets'))
    pan, ctr = pp.get_panoptic_segmentation(sem, hm, off, [1], 1000, 64, 0, 0.1, 7)
    pan2, ctr2 = pp.get_panoptic_segmentation(sem, hm, off, [1], 1000, 64, 0, 0.1, 7)
    assert torch.equal(pan, pan2) and torch.equal(ctr, ctr2)
    K = ctr.shape[1]
    assert 3000 < K <= 5000
    labs = torch.unique(pan)
    things = labs[labs > 0]
    assert int(things.min()) == 1001 and int(things.max()) == 1000 + things.numel() and things.numel() <= K
    assert bool(((pan[0, 0] > 0) == (sem[0, 0] == 1)).all())           # K > 20: ids are 0 only beyond 1e5 px
    ins, ctr3 = pp.get_instance_segmentation(sem[0], hm, off, [1], 0.1, 7)
    assert torch.equal(ctr3, ctr)
    merged = pp.merge_semantic_and_instance(sem[0], ins, 1000, [1], 64, 0)
    assert torch.equal(merged.reshape(pan.shape), pan)
    # centers are exactly the oracle's (cheap on the CPU even at full size)
    np.testing.assert_array_equal(ctr[0].cpu().numpy(), oracle.find_instance_center(d['ctr_hmp'][0, 0], 0.1, 7))


def test_dense_full_size_oracle(cuda_device):
    """BASELINE config 5 at full size, bit for bit: one 4096^2 tile with ~5000 small instances against the oracle
    (its masked brute-force search meets every one of the ~5000 centers at every thing pixel: ~10^10 distances)."""
    H = W = 4096
    d = synth_tile(H, W, 5000, seed=502, semi_axes=(4, 12), sigma=2.0)
    want_pan, want_ctr = oracle.get_panoptic_segmentation(d['sem'], d['ctr_hmp'], d['offsets'], [1], 1000, 64, 0, 0.1, 7)
    pan, ctr = pp.get_panoptic_segmentation(*(cu(d[k], cuda_device) for k in ('sem', 'ctr_hmp', 'offsets')), [1], 1000, 64, 0, 0.1, 7)
    np.testing.assert_array_equal(ctr.cpu().numpy(), want_ctr)
    assert ctr.shape[1] > 3000
    np.testing.assert_array_equal(pan.cpu().numpy(), want_pan)


def test_host_buffer_u8_entry_and_concurrent_callers(cuda_device):
    """emp_panoptic_batched_host_u8 (class maps already bytes on the host) gives the int64 entry's result, and two host
    threads calling the entry at once (the pipe serves one caller at a time) both get theirs."""
    import ctypes
    import threading
    from empanada_b200 import _cabi as C
    H, W, B = 192, 256, 4
    tiles = [synth_tile(H, W, 20 + 5 * i, seed=800 + i, semi_axes=(5, 16), sigma=3.0) for i in range(B)]
    sem_h = torch.from_numpy(np.stack([t['sem'][0, 0] for t in tiles])).pin_memory()
    sem8_h = sem_h.to(torch.uint8).pin_memory()
    hm_h = torch.from_numpy(np.stack([t['ctr_hmp'][0, 0] for t in tiles])).pin_memory()
    off_h = torch.from_numpy(np.stack([t['offsets'][0] for t in tiles])).pin_memory()
    L = C.lib()
    things, nt = C.i64_array([1])
    k_cap = 4096
    nbytes = L.emp_host_scratch_bytes(H, W, k_cap, nt)
    want = [oracle.get_panoptic_segmentation(t['sem'], t['ctr_hmp'], t['offsets'], [1], 1000, 64, 0, 0.1, 7)[0][0, 0] for t in tiles]
    results, errors = {}, []

    def call(name, fn, sem):
        try:
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=cuda_device)
            pan_h = torch.empty((B, H, W), dtype=torch.int64).pin_memory()
            k_out, f_out = (ctypes.c_int32 * B)(), (ctypes.c_int32 * B)()
            with torch.cuda.device(cuda_device):
                for _ in range(3):
                    C.check(fn(B, sem.data_ptr(), hm_h.data_ptr(), off_h.data_ptr(), H, W, things, nt, 1000, 64, 0, 0.1, 7,
                               pan_h.data_ptr(), k_out, f_out, k_cap, scratch.data_ptr(), nbytes))
            results[name] = pan_h.numpy().copy()
        except Exception as e:                      # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=call, args=('i64', L.emp_panoptic_batched_host, sem_h)),
               threading.Thread(target=call, args=('u8', L.emp_panoptic_batched_host_u8, sem8_h))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for name in ('i64', 'u8'):
        for b in range(B):
            np.testing.assert_array_equal(results[name][b], want[b])


def test_host_buffer_entry_point(cuda_device):
    """emp_panoptic_batched_host — the end-to-end entry bench.py times: pageable AND pinned host buffers
    in, H2D / kernels / D2H pipelined over three slots, results identical to the resident path and
    to the oracle, K and flags reported per tile."""
    import ctypes
    from empanada_b200 import _cabi as C
    H, W, B = 256, 320, 5
    tiles = [synth_tile(H, W, 30 + 7 * i, seed=700 + i, semi_axes=(5, 16), sigma=3.0) for i in range(B)]
    sem_h = torch.from_numpy(np.stack([t['sem'][0, 0] for t in tiles]))
    hm_h = torch.from_numpy(np.stack([t['ctr_hmp'][0, 0] for t in tiles]))
    off_h = torch.from_numpy(np.stack([t['offsets'][0] for t in tiles]))
    L = C.lib()
    things, nt = C.i64_array([1])
    k_cap = 4096
    nbytes = L.emp_host_scratch_bytes(H, W, k_cap, nt)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=cuda_device)
    for pin in (False, True):
        bufs = [x.pin_memory() if pin else x.clone() for x in (sem_h, hm_h, off_h)]
        pan_h = torch.empty((B, H, W), dtype=torch.int64)
        if pin:
            pan_h = pan_h.pin_memory()
        k_out, f_out = (ctypes.c_int32 * B)(), (ctypes.c_int32 * B)()
        with torch.cuda.device(cuda_device):
            C.check(L.emp_panoptic_batched_host(B, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), H, W, things, nt,
                                                1000, 64, 0, 0.1, 7, pan_h.data_ptr(), k_out, f_out, k_cap,
                                                scratch.data_ptr(), nbytes))
        for b, t in enumerate(tiles):
            want_pan, want_ctr = oracle.get_panoptic_segmentation(t['sem'], t['ctr_hmp'], t['offsets'], [1], 1000, 64, 0, 0.1, 7)
            assert k_out[b] == want_ctr.shape[1] and f_out[b] == 0
            np.testing.assert_array_equal(pan_h[b].numpy(), want_pan[0, 0])
        assert L.emp_host_sem_bytes_per_px() == 1.0                 # every class map crossed the link as bytes


def test_host_buffer_entry_wide_class_ids(cuda_device):
    """The host entry narrows int64 class maps to bytes before the PCIe copy; a tile holding a class id above 255
    (here a stuff class 300 and 256) or a negative id must travel as int64 and still give the oracle's answer /
    the class-range flag.  Odd sizes exercise the unaligned head / tail of the packing loop."""
    import ctypes
    from empanada_b200 import _cabi as C
    H, W, B = 131, 203, 6
    tiles = [synth_tile(H, W, 12 + 3 * i, seed=900 + i, semi_axes=(5, 14), sigma=3.0, stuff_classes=(2,)) for i in range(B)]
    tiles[1]['sem'][0, 0, 5:40, 7:90][tiles[1]['sem'][0, 0, 5:40, 7:90] == 0] = 300
    tiles[3]['sem'][0, 0, -1, -1] = 256                               # one pixel, the very last one
    tiles[4]['sem'][0, 0, 0, 0] = 255                                 # still a byte
    tiles[5]['sem'][0, 0, 64, 100] = -3                               # out of range: flagged, not narrowed
    sem_h = torch.from_numpy(np.stack([t['sem'][0, 0] for t in tiles])).pin_memory()
    hm_h = torch.from_numpy(np.stack([t['ctr_hmp'][0, 0] for t in tiles])).pin_memory()
    off_h = torch.from_numpy(np.stack([t['offsets'][0] for t in tiles])).pin_memory()
    pan_h = torch.empty((B, H, W), dtype=torch.int64).pin_memory()
    L = C.lib()
    things, nt = C.i64_array([1])
    k_cap = 2048
    nbytes = L.emp_host_scratch_bytes(H, W, k_cap, nt)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=cuda_device)
    k_out, f_out = (ctypes.c_int32 * B)(), (ctypes.c_int32 * B)()
    with torch.cuda.device(cuda_device):
        C.check(L.emp_panoptic_batched_host(B, sem_h.data_ptr(), hm_h.data_ptr(), off_h.data_ptr(), H, W, things, nt,
                                            1000, 20, 0, 0.1, 7, pan_h.data_ptr(), k_out, f_out, k_cap, scratch.data_ptr(), nbytes))
    for b, t in enumerate(tiles[:5]):
        want_pan, want_ctr = oracle.get_panoptic_segmentation(t['sem'], t['ctr_hmp'], t['offsets'], [1], 1000, 20, 0, 0.1, 7)
        assert k_out[b] == want_ctr.shape[1] and f_out[b] == 0
        np.testing.assert_array_equal(pan_h[b].numpy(), want_pan[0, 0])
    assert f_out[5] & C.FLAG_CLASS_RANGE
    assert L.emp_host_sem_bytes_per_px() == (3 * 1.0 + 3 * 8.0) / 6          # tiles 1, 3, 5 went as int64


def test_status_flags_through_the_c_abi(cuda_device):
    """Data-dependent conditions come back through the status block: more centers than k_cap
    (EMP_FLAG_K_OVERFLOW + true K) and class ids outside [0, 4096) (EMP_FLAG_CLASS_RANGE -> ValueError)."""
    from empanada_b200 import _cabi as C
    H, W = 64, 128
    hm = torch.zeros((1, 1, H, W), device=cuda_device)
    hm[0, 0, ::8, ::8] = 1.0                                        # 8 x 16 = 128 isolated peaks
    off = torch.zeros((1, 2, H, W), device=cuda_device)
    sem = torch.ones((1, 1, H, W), dtype=torch.int64, device=cuda_device)
    pan, ctr, Ks = pp._panoptic_tiles(sem.reshape(1, H, W), hm.reshape(1, H, W), off, [1], 1000, 0, 0, 0.1, 3, k_cap=16)
    assert Ks == [128] and ctr.shape[1] >= 128                      # retried with the true K
    want_pan, want_ctr = oracle.get_panoptic_segmentation(sem.cpu().numpy(), hm.cpu().numpy(), off.cpu().numpy(), [1], 1000, 0, 0, 0.1, 3)
    np.testing.assert_array_equal(pan.cpu().numpy()[0], want_pan[0, 0])
    bad = sem.clone()
    bad[0, 0, 5, 5] = 5000
    with pytest.raises(ValueError, match='class ids'):
        pp.get_panoptic_segmentation(bad, hm, off, [1], 1000, 0, 0, 0.1, 3)
    bad[0, 0, 5, 5] = -3
    with pytest.raises(ValueError, match='class ids'):
        pp.get_panoptic_segmentation(bad, hm, off, [1], 1000, 0, 0, 0.1, 3)
    assert C.lib().emp_version() >= 100
