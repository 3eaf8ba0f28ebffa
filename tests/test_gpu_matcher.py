"""B200: the cross-slice RLE matcher against fixtures produced by the reference's own
matcher.py (tests/golden/matcher_*.npz, incl. the reference's tests/test_matcher.py known answer):
the dict API (one overlap launch per call) and the block API (one launch per z-block)."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from empanada_b200.inference import matcher as mt
from empanada_b200.inference import rle

pytestmark = pytest.mark.gpu


def _unflatten(inst, starts, runs):
    seg, at = {}, 0
    for cls, lab, y0, x0, y1, x1, n in inst.tolist():
        seg[int(lab)] = {'box': (y0, x0, y1, x1), 'starts': starts[at:at + n], 'runs': runs[at:at + n]}
        at += n
    return seg


def _same(got, want):
    assert [int(k) for k in got.keys()] == list(want.keys())
    for k, a in want.items():
        g = got[k]
        assert tuple(int(v) for v in g['box']) == a['box']
        np.testing.assert_array_equal(np.asarray(g['starts']), a['starts'])
        np.testing.assert_array_equal(np.asarray(g['runs']), a['runs'])


def test_known_answer(cuda_device):
    """tests/test_matcher.py of the reference, plus the raw matrices of rle_matcher."""
    g = load_golden('matcher_known_answer')
    t = rle.pan_seg_to_rle_seg(g['target'], [1], 1000, [1], False)[1]
    m = rle.pan_seg_to_rle_seg(g['match'], [1], 1000, [1], False)[1]
    (a, b), all_labels, ious, iou, ioa = mt.rle_matcher(t, m, 0.25, return_iou=True, return_ioa=True)
    np.testing.assert_array_equal(a, g['matched_t'])
    np.testing.assert_array_equal(b, g['matched_m'])
    np.testing.assert_array_equal(ious, g['matched_ious'])
    assert iou.dtype == np.float64 and ioa.dtype == np.float32
    np.testing.assert_array_equal(iou, g['iou'])
    np.testing.assert_array_equal(ioa, g['ioa'])
    mat = mt.RLEMatcher(1, 1000, 0.25, 0.25, True)
    mat.initialize_target(t)
    out = mat(m, update_target=False)
    np.testing.assert_array_equal(rle.rle_seg_to_pan_seg({1: out}, (200, 200)), g['out'])
    # empty sides
    (e1, e2), (l1, l2), e3, e4 = mt.rle_matcher({}, m, 0.25, return_ioa=True)
    assert e1.size == 0 and l1.size == 0 and len(l2) == len(m) and e4.size == 0


@pytest.mark.parametrize('name', golden_names('matcher_stack_'))
@pytest.mark.parametrize('api', ['dict', 'block'])
def test_forward_backward_stack(name, api, cuda_device):
    g = load_golden(name)
    p = g['params']
    D, H, W = p['D'], p['H'], p['W']
    vol = torch.from_numpy(g['in_vol']).to(cuda_device)
    if api == 'dict':
        rles = [rle.pan_seg_to_rle_seg(vol[z], [1], 1000, [1], p['force_connected'])[1] for z in range(D)]
        mat = mt.RLEMatcher(1, 1000, 0.25, 0.25, True)
        fwd = []
        for z in range(D):
            seg = rles[z]
            if mat.target_rle is None:
                mat.initialize_target(seg)
            else:
                seg = mat(seg)
            fwd.append(seg)
        assert mat.next_label == int(g['fwd_next_label'])
        mat.target_rle, mat.assign_new = None, False
        bwd = [None] * D
        for z in range(D - 1, -1, -1):
            seg = fwd[z]
            if mat.target_rle is None:
                mat.initialize_target(seg)
            else:
                seg = mat(seg)
            bwd[z] = seg
    else:
        run_cap = 1 << 14
        runs_all = torch.empty((D, run_cap, 3), dtype=torch.int64, device=cuda_device)
        inst_all = torch.empty((D, run_cap, 8), dtype=torch.int64, device=cuda_device)
        rles, n_runs = [], []
        for z in range(D):
            ws = rle.rle_enqueue(vol[z].contiguous(), [1], 1000, [1], p['force_connected'], runs_all[z], inst_all[z])
            st = ws[:64].view(torch.int32).cpu()
            nr, ni = int(st[3]), int(st[4])
            n_runs.append(nr)
            rles.append(rle.tables_to_rle_seg(inst_all[z, :ni].cpu().numpy(), runs_all[z, :nr].cpu().numpy(), [1])[1])
        overlaps = mt.block_overlaps(runs_all, n_runs)
        assert len(overlaps) == D - 1
        sm = mt.StackMatcher(1, 1000, 0.25, 0.25)
        fwd, groups = sm.forward(rles, overlaps)
        assert sm.matcher.next_label == int(g['fwd_next_label'])
        bwd = sm.backward(fwd, groups, rles, overlaps)
        # dense fill of the block straight from the run tables + the matcher's per-slot labels
        from empanada_b200.inference import fill as fl
        width = max(len(l) for l in sm.slot_labels)
        table = np.full((D, max(width, 1)), -1, np.int64)
        for z in range(D):
            table[z, :len(sm.slot_labels[z])] = sm.slot_labels[z]
        for dt in (torch.int64, torch.int32):
            vol_out = fl.fill_block(runs_all, n_runs, table, (H, W), dt).cpu().numpy()
            for z in range(D):
                np.testing.assert_array_equal(vol_out[z], g[f'bwd_{z}'])
    for z in range(D):
        _same(fwd[z], _unflatten(g[f'fwd_inst_{z}'], g[f'fwd_starts_{z}'], g[f'fwd_runs_{z}']))
        _same(bwd[z], _unflatten(g[f'bwd_inst_{z}'], g[f'bwd_starts_{z}'], g[f'bwd_runs_{z}']))
        np.testing.assert_array_equal(rle.rle_seg_to_pan_seg({1: bwd[z]}, (H, W)), g[f'bwd_{z}'])


def test_fill_instances_dict_api(cuda_device):
    """fill_instances on a flat 3D volume, incl. overlapping instances (later ones win, as in the
    reference's loop) and untouched voxels keeping their value."""
    from empanada_b200.inference import fill as fl
    rng = np.random.default_rng(3)
    shape = (5, 40, 64)
    n = int(np.prod(shape))
    inst = {}
    for lab in (7, 1003, 20002, 5):
        starts = np.sort(rng.choice(n - 50, 30, replace=False)).astype(np.int64)
        inst[lab] = {'box': None, 'starts': starts, 'runs': rng.integers(1, 40, 30).astype(np.int64)}
    want = np.full(n, 9, np.int64)
    for lab, a in inst.items():
        for s0, r0 in zip(a['starts'], a['runs']):
            want[s0:s0 + r0] = lab
    vol = torch.full(shape, 9, dtype=torch.int64, device=cuda_device)
    got = fl.fill_instances(vol, inst)
    np.testing.assert_array_equal(got.cpu().numpy().ravel(), want)
    assert fl.fill_instances(torch.zeros(4, 4, dtype=torch.int32, device=cuda_device), {}).sum() == 0


def test_merge_rles_host_semantics():
    """Touching and overlapping ranges join; output sorted (array_utils.py:634-718)."""
    s, r = mt.merge_rles(np.array([0, 10, 30]), np.array([5, 5, 5]), np.array([5, 14, 50]), np.array([5, 3, 2]))
    np.testing.assert_array_equal(s, [0, 30, 50])
    np.testing.assert_array_equal(r, [17, 5, 2])
    assert mt.merge_boxes((1, 5, 9, 9), (0, 6, 4, 12)) == (0, 5, 9, 12)


def test_pair_overlaps_with_overlapping_instances(cuda_device):
    """The dict API takes any instance lists: instances that overlap each other (merged trackers, hand-made dicts) must
    still give the reference's per-pair intersections (array_utils.rle_intersection :371-403 treats every pair alone)."""
    from oracle import matcher as om
    rng = np.random.default_rng(8)

    def inst(n):
        out = []
        for _ in range(n):
            s = np.sort(rng.choice(4000, size=9, replace=False)).astype(np.int64)
            r = np.minimum(rng.integers(1, 120, size=9), np.append(np.diff(s), 200)).astype(np.int64)   # disjoint inside one instance
            out.append((s, r))
        return out
    A, B = inst(7), inst(5)                         # instances of A overlap each other freely, so do B's
    want = np.array([[om.rle_intersection(sa, ra, sb, rb) for sb, rb in B] for sa, ra in A], np.int64)
    got = mt.pair_overlaps([s for s, _ in A], [r for _, r in A], [s for s, _ in B], [r for _, r in B], cuda_device)
    np.testing.assert_array_equal(got, want)
    assert (want > 0).sum() > 10
