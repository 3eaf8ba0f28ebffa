"""CPU: the oracle (oracle/oracle.c + oracle/__init__.py) against fixtures produced by running
the reference itself (tests/golden/make_golden.py).  Bit-exact everywhere."""
import numpy as np
import pytest

import oracle
from conftest import golden_names, load_golden


@pytest.mark.parametrize('name', golden_names('pp_'))
def test_postprocess_matches_reference(name):
    g = load_golden(name)
    p = g['params']
    ctr = oracle.find_instance_center(g['in_hm'], p['threshold'], p['nms_kernel'])
    np.testing.assert_array_equal(ctr, g['out_centers'])
    for step in (1, 4):
        key = f'out_ids_step{step}'
        if key in g:
            ids = oracle.group_pixels(ctr, g['in_off'], step=float(step))
            np.testing.assert_array_equal(ids, g[key])
    ins, c = oracle.get_instance_segmentation(g['in_sem'], g['in_hm'], g['in_off'], p['thing_list'],
                                              p['threshold'], p['nms_kernel'])
    np.testing.assert_array_equal(ins, g['out_ins'])
    pan, c = oracle.get_panoptic_segmentation(g['in_sem'], g['in_hm'], g['in_off'], p['thing_list'],
                                              p['label_divisor'], p['stuff_area'], p['void_label'],
                                              p['threshold'], p['nms_kernel'])
    assert pan.shape == g['out_pan'].shape and pan.dtype == np.int64
    np.testing.assert_array_equal(pan, g['out_pan'])


@pytest.mark.parametrize('name', golden_names('merge_'))
def test_merge_matches_reference(name):
    g = load_golden(name)
    p = g['params']
    pan = oracle.merge_semantic_and_instance(g['in_sem'], g['in_ins'], p['label_divisor'],
                                             p['thing_list'], p['stuff_area'], p['void_label'])
    assert pan.shape == g['out_pan'].shape
    np.testing.assert_array_equal(pan, g['out_pan'])


def _sigmoid_or_softmax(logits):
    import torch          # engines.logits_to_prob (engines.py:22-30) stays a torch op in the product too
    t = torch.from_numpy(logits)
    return (torch.softmax(t, 1) if t.size(1) > 1 else torch.sigmoid(t)).numpy()


@pytest.mark.parametrize('name', golden_names('engine3d_'))
def test_median_queue_and_engine3d(name):
    g = load_golden(name)
    p = g['params']
    q = oracle.MedianQueue(p['median_kernel_size'])
    outs = []

    def post(entry):
        sem = oracle.harden_seg(entry['sem'], p['confidence_thr'])
        pan, _ = oracle.get_panoptic_segmentation(sem, entry['ctr_hmp'], entry['offsets'], p['thing_list'],
                                                  p['label_divisor'], p['stuff_area'], p['void_label'],
                                                  p['nms_threshold'], p['nms_kernel'])
        return pan

    emitted = []
    for z in range(p['n']):
        e = {'sem': _sigmoid_or_softmax(g[f'in_{z}_sem_logits']), 'ctr_hmp': g[f'in_{z}_ctr_hmp'],
             'offsets': g[f'in_{z}_offsets']}
        q.enqueue(e)
        o = q.get_next(['sem'])
        emitted.append(o is not None)
        if o is not None:
            outs.append(post(o))
    for e in q.end():
        outs.append(post(e))
    assert emitted == p['emitted']
    assert len(outs) == sum(p['emitted']) + p['n_tail']
    for i, o in enumerate(outs):
        np.testing.assert_array_equal(o, g[f'out_{i}'])


@pytest.mark.parametrize('name', golden_names('render3d_'))
def test_render_engine3d(name):
    g = load_golden(name)
    p = g['params']
    q = oracle.MedianQueue(p['median_kernel_size'])
    step = 4 if p['coarse_boundaries'] else 1
    up = p['upsampling']
    h, w = p['size']
    outs = []

    def post(entry):
        ctr = oracle.find_instance_center(entry['ctr_hmp'], p['nms_threshold'], p['nms_kernel'])
        if ctr.shape[0] == 0:
            cells = np.zeros(entry['ctr_hmp'].shape[-2:], np.int64)
        else:
            cells = oracle.group_pixels(ctr, entry['offsets'], step=float(step))[0]
        cells = oracle.nearest_upsample(cells, int(up * step))
        sem = oracle.harden_seg(entry['sem'], p['confidence_thr'])[0]       # (1,H,W)
        thing = np.isin(sem, p['thing_list'])
        ins = np.where(thing, cells[None], 0)
        pan = oracle.merge_semantic_and_instance(sem, ins, p['label_divisor'], p['thing_list'],
                                                 p['stuff_area'], p['void_label'])
        return pan[..., :h, :w]

    for z in range(p['n']):
        e = {'sem': _sigmoid_or_softmax(g[f'in_{z}_sem_logits']), 'ctr_hmp': g[f'in_{z}_ctr_hmp'],
             'offsets': g[f'in_{z}_offsets']}
        q.enqueue(e)
        o = q.get_next(['sem'])
        if o is not None:
            outs.append(post(o))
    for e in q.end():
        outs.append(post(e))
    assert len(outs) == sum(p['emitted']) + p['n_tail']
    for i, o in enumerate(outs):
        assert o.shape == g[f'out_{i}'].shape
        np.testing.assert_array_equal(o, g[f'out_{i}'])


def _unflatten_labels(inst, starts, runs, labels):
    seg, at = {int(l): {} for l in labels}, 0
    for row in inst:
        n = int(row[6])
        seg[int(row[0])][int(row[1])] = {'box': tuple(int(v) for v in row[2:6]), 'starts': starts[at:at + n], 'runs': runs[at:at + n]}
        at += n
    return seg


@pytest.mark.parametrize('name', golden_names('stack_rle_'))
def test_stack_of_rle_matches_reference(name):
    """The whole stack path — recursive median queue, harden, coarse cells, merge, crop, pan_seg_to_rle_seg with
    force_connected — restated by the oracle against the reference's own RLE tables for every slice."""
    g = load_golden(name)
    p = g['params']
    q = oracle.MedianQueue(p['median_kernel_size'])
    up = p['upsampling']
    h, w = p['size']
    outs = []

    def post(entry):
        ctr = oracle.find_instance_center(entry['ctr_hmp'], p['nms_threshold'], p['nms_kernel'])
        cells = np.zeros(entry['ctr_hmp'].shape[-2:], np.int64) if ctr.shape[0] == 0 else oracle.group_pixels(ctr, entry['offsets'], step=4.0)[0]
        cells = oracle.nearest_upsample(cells, int(up * 4))
        sem = oracle.harden_seg(entry['sem'], p['confidence_thr'])[0]
        ins = np.where(np.isin(sem, p['thing_list']), cells[None], 0)
        pan = oracle.merge_semantic_and_instance(sem, ins, p['label_divisor'], p['thing_list'], p['stuff_area'], p['void_label'])
        return oracle.pan_seg_to_rle_seg(pan[0, :h, :w], p['labels'], p['label_divisor'], p['thing_list'], True)

    for z in range(p['n']):
        q.enqueue({'sem': _sigmoid_or_softmax(g[f'in_{z}_sem_logits']), 'ctr_hmp': g[f'in_{z}_ctr_hmp'], 'offsets': g[f'in_{z}_offsets']})
        o = q.get_next(['sem'])
        if o is not None:
            outs.append(post(o))
    outs += [post(e) for e in q.end()]
    assert len(outs) == p['n']
    for z, got in enumerate(outs):
        want = _unflatten_labels(g[f"out_{z}_inst"], g[f"out_{z}_starts"], g[f"out_{z}_runs"], p["labels"])
        assert list(got.keys()) == list(want.keys())
        for c in want:
            assert list(got[c].keys()) == list(want[c].keys()), f'slice {z} class {c}'
            for lab in want[c]:
                assert tuple(got[c][lab]['box']) == want[c][lab]['box']
                np.testing.assert_array_equal(got[c][lab]['starts'], want[c][lab]['starts'])
                np.testing.assert_array_equal(got[c][lab]['runs'], want[c][lab]['runs'])


@pytest.mark.parametrize('name', golden_names('rle_'))
def test_rle_matches_reference(name):
    g = load_golden(name)
    p = g['params']
    seg = oracle.pan_seg_to_rle_seg(g['in_pan'], p['labels'], p['label_divisor'], p['thing_list'],
                                    p['force_connected'])
    inst, starts, runs = [], [], []
    assert list(seg.keys()) == p['labels']
    for cls, attrs in seg.items():
        for lab, a in attrs.items():
            inst.append([cls, lab, *a['box'], len(a['starts'])])
            starts += list(a['starts'])
            runs += list(a['runs'])
    np.testing.assert_array_equal(np.asarray(inst, np.int64).reshape(-1, 7), g['out_inst'])
    np.testing.assert_array_equal(np.asarray(starts, np.int64), g['out_starts'])
    np.testing.assert_array_equal(np.asarray(runs, np.int64), g['out_runs'])
    np.testing.assert_array_equal(oracle.rle_seg_to_pan_seg(seg, g['in_pan'].shape), g['out_back'])


def test_sigmoid_cpu_matches_goldens_engine2d():
    g = load_golden('engine2d')
    p = g['params']
    for z in range(p['n']):
        sem = oracle.harden_seg(_sigmoid_or_softmax(g[f'in_{z}_sem_logits']), p['confidence_thr'])
        pan, _ = oracle.get_panoptic_segmentation(sem, g[f'in_{z}_ctr_hmp'], g[f'in_{z}_offsets'],
                                                  p['thing_list'], p['label_divisor'], p['stuff_area'],
                                                  p['void_label'], p['nms_threshold'], p['nms_kernel'])
        np.testing.assert_array_equal(pan, g[f'out_{z}'])


@pytest.mark.parametrize('name', ['pp_fixture256', 'pp_chunked_k_gt_20', 'pp_ties_half_offsets', 'pp_multiclass_1',
                                  'pp_sentinel_k21', 'pp_plateau_k4', 'pp_k0'])
def test_torch_port_matches_reference(name):
    """The op-sequence port that bench.py times as the CPU baseline gives the reference's results."""
    import torch
    from oracle import torch_port
    g = load_golden(name)
    p = g['params']
    pan, ctr = torch_port.panoptic(torch.from_numpy(g['in_sem']), torch.from_numpy(g['in_hm']),
                                   torch.from_numpy(g['in_off']), p['thing_list'], p['label_divisor'],
                                   p['stuff_area'], p['void_label'], p['threshold'], p['nms_kernel'])
    np.testing.assert_array_equal(ctr[0].numpy(), g['out_centers'])
    np.testing.assert_array_equal(pan.numpy(), g['out_pan'])


# ---- cross-slice matcher -----------------------------------------------------------
def _unflatten(inst, starts, runs):
    seg, at = {}, 0
    for cls, lab, y0, x0, y1, x1, n in inst.tolist():
        seg.setdefault(int(cls), {})[int(lab)] = {'box': (y0, x0, y1, x1), 'starts': starts[at:at + n], 'runs': runs[at:at + n]}
        at += n
    return seg


def _same_rles(got, inst, starts, runs):
    want = _unflatten(inst, starts, runs).get(1, {})
    assert [int(k) for k in got.keys()] == list(want.keys())
    for k, a in want.items():
        g = got[k]
        assert tuple(int(v) for v in g['box']) == a['box']
        np.testing.assert_array_equal(np.asarray(g['starts']), a['starts'])
        np.testing.assert_array_equal(np.asarray(g['runs']), a['runs'])


def test_matcher_oracle_known_answer():
    from oracle import matcher as om
    g = load_golden('matcher_known_answer')
    t = oracle.pan_seg_to_rle_seg(g['target'], [1], 1000, [1], False)[1]
    m = oracle.pan_seg_to_rle_seg(g['match'], [1], 1000, [1], False)[1]
    (mt, mm), _, ious, iou, ioa = om.rle_matcher(t, m, 0.25, return_iou=True, return_ioa=True)
    np.testing.assert_array_equal(mt, g['matched_t'])
    np.testing.assert_array_equal(mm, g['matched_m'])
    np.testing.assert_array_equal(ious, g['matched_ious'])
    np.testing.assert_array_equal(iou, g['iou'])
    np.testing.assert_array_equal(ioa, g['ioa'])
    mat = om.RLEMatcher(1, 1000, 0.25, 0.25, True)
    mat.initialize_target(t)
    out = mat(m, update_target=False)
    np.testing.assert_array_equal(oracle.rle_seg_to_pan_seg({1: out}, (200, 200)), g['out'])


@pytest.mark.parametrize('name', golden_names('matcher_stack_'))
def test_matcher_oracle_stack(name):
    from oracle import matcher as om
    g = load_golden(name)
    p = g['params']
    D = p['D']
    rles = [oracle.pan_seg_to_rle_seg(g['in_vol'][z], [1], 1000, [1], p['force_connected'])[1] for z in range(D)]
    mat = om.RLEMatcher(1, 1000, 0.25, 0.25, True)
    fwd = []
    for z in range(D):
        seg = rles[z]
        if mat.target_rle is None:
            mat.initialize_target(seg)
        else:
            seg = mat(seg)
        fwd.append(seg)
        _same_rles(seg, g[f'fwd_inst_{z}'], g[f'fwd_starts_{z}'], g[f'fwd_runs_{z}'])
    assert mat.next_label == int(g['fwd_next_label'])
    mat.target_rle, mat.assign_new = None, False
    for z in range(D - 1, -1, -1):
        seg = fwd[z]
        if mat.target_rle is None:
            mat.initialize_target(seg)
        else:
            seg = mat(seg)
        _same_rles(seg, g[f'bwd_inst_{z}'], g[f'bwd_starts_{z}'], g[f'bwd_runs_{z}'])


def test_config1_matches_reference_run():
    """BASELINE configs[0]: the oracle on the 1024^2 tile the reference's own engine (ResNet-50 PDL + post-processing,
    tests/golden/make_config1.py) was run on."""
    import torch
    from empanada_b200.synth import CONFIG1 as p, config1_heads
    g = load_golden('config1')
    h = config1_heads(g['consts'])
    sem = oracle.harden_seg(torch.sigmoid(torch.from_numpy(h['sem_logits'])).numpy(), p['confidence_thr'])
    pan, _ = oracle.get_panoptic_segmentation(sem, h['ctr_hmp'], h['offsets'], p['thing_list'], p['label_divisor'],
                                              p['stuff_area'], p['void_label'], p['nms_threshold'], p['nms_kernel'])
    assert np.unique(g['pan']).size > 50
    np.testing.assert_array_equal(pan.reshape(g['pan'].shape), g['pan'])
