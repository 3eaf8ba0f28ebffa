"""
make_golden.py — run the REFERENCE ITSELF (imported unmodified from /root/reference) on small
seeded inputs and store inputs + outputs as fixtures under tests/golden/.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

What is executed is the reference's own code: empanada/inference/postprocess.py and engines.py
as they are; empanada/inference/rle.py through the skimage/zarr shim of oracle/ref_shim.py;
data/utils/target_creation.py (loaded by path) for the 256x256 test fixture of the reference's
tests/test_data_post.py.  The tests (tests/test_oracle_golden.py on CPU, tests/test_gpu_*.py on
the B200) compare the oracle and the CUDA path with these files bit for bit.
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.install()

import torch  # noqa: E402
from empanada.inference import postprocess as rpp  # noqa: E402  (reference)
from empanada.inference import engines as reng  # noqa: E402  (reference)
from empanada.inference import rle as rrle  # noqa: E402  (reference, via shim)
from empanada_b200.synth import synth_tile  # noqa: E402

torch.set_num_threads(8)


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name}: {os.path.getsize(path) / 1024:.0f} KiB')


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


# ---------------------------------------------------------------------------------------------
# post-processing cases
# ---------------------------------------------------------------------------------------------
def run_pp(name, sem, hm, off, thing_list, L, stuff_area, void, thr=0.1, k=7, steps=(1.0, 4.0)):
    sem, hm, off = (np.ascontiguousarray(a) for a in (sem, hm, off))
    out = {'in_sem': sem, 'in_hm': hm, 'in_off': off,
           'params': json.dumps(dict(thing_list=list(thing_list), label_divisor=L,
                                     stuff_area=stuff_area, void_label=void,
                                     threshold=thr, nms_kernel=k))}
    ctr = rpp.find_instance_center(t(hm), thr, k)
    out['out_centers'] = ctr.numpy()
    if ctr.size(0) > 0:
        for s in steps:
            out[f'out_ids_step{int(s)}'] = rpp.group_pixels(ctr, t(off), step=float(s)).numpy()
    ins, c2 = rpp.get_instance_segmentation(t(sem), t(hm), t(off), list(thing_list), thr, k)
    out['out_ins'] = ins.numpy()
    pan, c3 = rpp.get_panoptic_segmentation(t(sem), t(hm), t(off), list(thing_list), L,
                                            stuff_area, void, thr, k)
    out['out_pan'] = pan.numpy()
    assert c3.shape[1] == ctr.shape[0]
    print(f'  {name}: K={ctr.size(0)} pan shape {tuple(pan.shape)}')
    save(name, **out)


def fixture256():
    """The reference's own test fixture (tests/test_data_post.py:13-36): GT mask ->
    sem / heat-map / offsets exactly as data/panoptic_dataset.py:72-95 builds them."""
    import cv2
    spec = importlib.util.spec_from_file_location(
        'ref_target_creation',
        os.path.join(ref_shim.REFERENCE_ROOT, 'empanada/data/utils/target_creation.py'))
    tc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tc)
    mask = cv2.imread(os.path.join(ref_shim.REFERENCE_ROOT,
                                   'tests/test_data/panoptic/dataset1/masks/pan_seg.tiff'),
                      cv2.IMREAD_UNCHANGED)
    labels, things, L = [1, 2, 3], [2], 1000
    thing_seg = np.zeros_like(mask)
    sem_seg = np.zeros_like(mask)
    for c in labels:
        inside = (mask >= c * L) * (mask < (c + 1) * L)
        sem_seg[inside] = c
        if c in things:
            thing_seg[inside] = mask[inside]
    hm, off = tc.heatmap_and_offsets(thing_seg, 6)
    sem = sem_seg.astype(np.int32).astype(np.int64)[None, None]
    run_pp('pp_fixture256', sem, hm[None], off[None], things, L, 0, 0, 0.1, 7)
    np.savez_compressed(os.path.join(HERE, 'fixture256_mask.npz'), mask=mask)


def pp_cases():
    fixture256()

    d = synth_tile(96, 128, 9, seed=1, semi_axes=(6, 14))
    run_pp('pp_small_k_le_20', d['sem'], d['ctr_hmp'], d['offsets'], [1], 1000, 64, 0)

    d = synth_tile(128, 160, 70, seed=2, semi_axes=(4, 9), sigma=2.0)
    run_pp('pp_chunked_k_gt_20', d['sem'], d['ctr_hmp'], d['offsets'], [1], 1000, 64, 0, 0.1, 3)

    # exact ties: zero offsets on a lattice of centers, and integer / half-integer offsets
    rng = np.random.default_rng(3)
    H, W = 64, 80
    hm = np.zeros((H, W), np.float32)
    hm[8::16, 8::16] = 1.0
    off = np.zeros((2, H, W), np.float32)
    sem = np.ones((1, 1, H, W), np.int64)
    run_pp('pp_ties_zero_offsets', sem, hm[None, None], off[None], [1], 1000, 0, 0, 0.1, 7)
    off = rng.integers(-12, 13, (2, H, W)).astype(np.float32)
    run_pp('pp_ties_int_offsets', sem, hm[None, None], off[None], [1], 1000, 0, 0, 0.1, 7)
    off = (rng.integers(-24, 25, (2, H, W)) * 0.5).astype(np.float32)
    hm2 = np.zeros((H, W), np.float32)
    hm2[4::8, 4::8] = 1.0                       # 80 centers -> chunked path with many ties
    run_pp('pp_ties_half_offsets', sem, hm2[None, None], off[None], [1], 1000, 0, 0, 0.1, 3)

    # plateaus / duplicated maxima, several kernels (even one included) and thresholds
    for k, thr in ((1, 0.3), (3, 0.1), (4, 0.5), (7, -0.2), (5, 0.1)):
        hm = (rng.integers(0, 6, (48, 56)) / 5.0).astype(np.float32)
        hm[10:14, 20:25] = 1.0
        hm[0, 0] = 1.0
        hm[-1, -1] = 1.0
        hm[30, :] = 0.8
        off = rng.normal(0, 3, (2, 48, 56)).astype(np.float32)
        sem = (rng.random((1, 1, 48, 56)) < 0.7).astype(np.int64)
        run_pp(f'pp_plateau_k{k}', sem, hm[None, None], off[None], [1], 1000, 16, 0, thr, k,
               steps=(1.0,))

    # 1e5 sentinel: K = 21 (> chunksize) and offsets pushing a region beyond 1e5 -> id 0
    H, W = 40, 64
    hm = np.zeros((H, W), np.float32)
    hm[4::8, 4::8][:3, :7] = 1.0              # 21 centers
    off = rng.normal(0, 2, (2, H, W)).astype(np.float32)
    off[:, 10:20, 10:30] += 2e5
    sem = np.ones((1, 1, H, W), np.int64)
    run_pp('pp_sentinel_k21', sem, hm[None, None], off[None], [1], 1000, 0, 0, 0.1, 7)
    hm[4::8, 4::8][:3, :7] = 0
    hm[4::8, 4::8][:2, :5] = 1.0              # 10 centers: no sentinel
    run_pp('pp_no_sentinel_k10', sem, hm[None, None], off[None], [1], 1000, 0, 0, 0.1, 7)

    # multi-class: 2 thing classes + 2 stuff classes, vote ties, void labels, stuff thresholds
    for i, (void, sa) in enumerate(((0, 0), (-1, 64), (7, 200))):
        d = synth_tile(96, 112, 30, seed=10 + i, semi_axes=(5, 12), sigma=2.5,
                       thing_classes=(1, 3), stuff_classes=(2, 4))
        sem = d['sem'].copy()
        # scramble classes inside instances so that votes are contested (incl. exact ties)
        noise = rng.integers(0, 5, sem.shape)
        flip = rng.random(sem.shape) < 0.45
        sem[flip] = noise[flip]
        sem[0, 0, 40:44, 40:60] = 1
        sem[0, 0, 44:48, 40:60] = 3
        run_pp(f'pp_multiclass_{i}', sem, d['ctr_hmp'], d['offsets'], [1, 3], 1000, sa, void,
               0.1, 5, steps=(1.0,))

    # no centers at all
    d = synth_tile(32, 48, 4, seed=20, semi_axes=(4, 8))
    run_pp('pp_k0', d['sem'], d['ctr_hmp'] * 0.05, d['offsets'], [1], 1000, 8, 0)

    # instances touching all borders, non-multiple-of-anything shape
    d = synth_tile(67, 93, 40, seed=21, semi_axes=(5, 15), sigma=3.0)
    run_pp('pp_odd_shape', d['sem'], d['ctr_hmp'], d['offsets'], [1], 20000, 32, 0, 0.1, 3)


def merge_cases():
    rng = np.random.default_rng(30)
    for i, (shape_sem, shape_ins) in enumerate((((1, 50, 60), (1, 50, 60)),
                                                ((1, 1, 50, 60), (1, 50, 60)),
                                                ((1, 37, 41), (1, 37, 41)))):
        H, W = shape_sem[-2:]
        sem = rng.integers(0, 5, shape_sem).astype(np.int64)
        ins = np.zeros((H, W), np.int64)
        for j in range(1, 40):
            y0, x0 = rng.integers(0, H - 4), rng.integers(0, W - 4)
            ins[y0:y0 + rng.integers(1, 9), x0:x0 + rng.integers(1, 9)] = j * (3 if i == 2 else 1)
        ins = ins.reshape(shape_ins)
        for void, sa, things in ((0, 64, [1, 2]), (-1, 0, [3]), (7, 200, [1, 2, 4])):
            pan = rpp.merge_semantic_and_instance(t(sem), t(ins), 1000, things, sa, void).numpy()
            save(f'merge_{i}_void{void}_sa{sa}', in_sem=sem, in_ins=ins, out_pan=pan,
                 params=json.dumps(dict(label_divisor=1000, thing_list=things, stuff_area=sa,
                                        void_label=void)))


# ---------------------------------------------------------------------------------------------
# engines: fake models that replay precomputed head tensors through the reference engines
# ---------------------------------------------------------------------------------------------
class ReplayModel(torch.nn.Module):
    def __init__(self, outputs):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.outputs = outputs
        self.i = 0

    def forward(self, image, render_steps=None, interpolate_ins=None):
        out = {k: v.clone() for k, v in self.outputs[self.i].items()}
        self.i += 1
        return out


def logit(p):
    return np.log(p / (1 - p)).astype(np.float32)


def stack_inputs(D, H, W, seed, coarse, n_classes=1):
    """Per-slice head tensors for a drifting blob field (small, deterministic)."""
    outs = []
    for z in range(D):
        d = synth_tile(H, W, 14, seed=seed, semi_axes=(5 + 0.3 * z, 11 + 0.3 * z), sigma=2.5,
                       prob=True)
        rng = np.random.default_rng([seed, z])
        prob = np.clip(d['sem_prob'] + rng.normal(0, 0.25, d['sem_prob'].shape), 0.02, 0.98).astype(np.float32)
        if n_classes == 1:
            sem_logits = logit(prob)
        else:
            sem_logits = rng.normal(0, 1, (1, n_classes, H, W)).astype(np.float32)
            sem_logits[:, 1] += 3 * (d['ins'] > 0)
        if coarse == 1:
            hm, off = d['ctr_hmp'], d['offsets']
        else:
            dd = synth_tile(H // coarse, W // coarse, 14, seed=seed, semi_axes=(2, 4), sigma=1.2)
            hm, off = dd['ctr_hmp'], dd['offsets'] * coarse
        outs.append({'sem_logits': sem_logits, 'ctr_hmp': hm.astype(np.float32), 'offsets': off.astype(np.float32)})
    return outs


def engine_cases():
    # 2D engine
    ins2d = stack_inputs(2, 64, 80, seed=40, coarse=1)
    model = ReplayModel([{k: t(v) for k, v in o.items()} for o in ins2d])
    eng = reng.PanopticDeepLabEngine(model, [1], 1000, 16, 0, 0.1, 3, 0.5)
    res = {}
    for z, o in enumerate(ins2d):
        res[f'out_{z}'] = eng(torch.zeros(1, 1, 64, 80)).numpy()
        for k, v in o.items():
            res[f'in_{z}_{k}'] = v
    save('engine2d', params=json.dumps(dict(thing_list=[1], label_divisor=1000, stuff_area=16,
                                            void_label=0, nms_threshold=0.1, nms_kernel=3,
                                            confidence_thr=0.5, n=2)), **res)

    # 3D engines: (name, class, ks, classes, coarse, upsampling, D)
    for name, ks, ncls, D in (('engine3d_ks3', 3, 1, 7), ('engine3d_ks5_mc', 5, 3, 8),
                              ('engine3d_ks1', 1, 1, 3), ('engine3d_short', 3, 1, 2)):
        ins = stack_inputs(D, 48, 64, seed=41, coarse=1, n_classes=ncls)
        model = ReplayModel([{k: t(v) for k, v in o.items()} for o in ins])
        things = [1] if ncls == 1 else [1, 2]
        eng = reng.PanopticDeepLabEngine3d(model, things, 1000, 16, 0, 0.1, 3, 0.3, ks)
        res, emitted = {}, []
        for z, o in enumerate(ins):
            for k, v in o.items():
                res[f'in_{z}_{k}'] = v
            out = eng(torch.zeros(1, 1, 48, 64))
            if out is not None:
                res[f'out_{sum(emitted)}'] = out.numpy()
            emitted.append(out is not None)
        tail = eng.end()
        n_call = sum(emitted)
        for j, out in enumerate(tail):
            res[f'out_{n_call + j}'] = out.numpy()
        save(name, params=json.dumps(dict(thing_list=things, label_divisor=1000, stuff_area=16,
                                          void_label=0, nms_threshold=0.1, nms_kernel=3,
                                          confidence_thr=0.3, median_kernel_size=ks, n=D,
                                          emitted=emitted, n_tail=len(tail))), **res)

    for name, ks, coarse, up, D, size in (('render3d_coarse', 3, 4, 1, 6, (60, 75)),
                                          ('render3d_fine_up2', 3, 1, 2, 5, (64, 80)),
                                          ('render3d_coarse_up2', 1, 4, 2, 3, (64, 80))):
        H, W = 64, 80                       # padded size (multiple of padding_factor 16)
        # model output res: sem at (H*up, W*up); ctr/offsets at (H*up/coarse... ) — the model
        # sees the padded image and render_steps; here we simply replay tensors of the shapes
        # the quantizable PR models emit (quantization/panoptic_deeplab.py:221-250)
        ins = []
        base = stack_inputs(D, H * up, W * up, seed=42, coarse=1)
        for z in range(D):
            o = dict(base[z])
            if coarse == 4:
                dd = synth_tile(H // 4, W // 4, 10, seed=43, semi_axes=(2, 4), sigma=1.2)
                o['ctr_hmp'] = dd['ctr_hmp'].astype(np.float32)
                o['offsets'] = (dd['offsets'] * 4).astype(np.float32)
            else:
                dd = synth_tile(H, W, 14, seed=43, semi_axes=(5, 11), sigma=2.5)
                o['ctr_hmp'] = dd['ctr_hmp'].astype(np.float32)
                o['offsets'] = dd['offsets'].astype(np.float32)
            ins.append(o)
        model = ReplayModel([{k: t(v) for k, v in o.items()} for o in ins])
        eng = reng.PanopticDeepLabRenderEngine3d(model, [1], 20000, 32, 0, 0.1, 3, 0.3, ks, 16,
                                                 coarse_boundaries=(coarse == 4))
        res, emitted = {}, []
        for z, o in enumerate(ins):
            for k, v in o.items():
                res[f'in_{z}_{k}'] = v
            out = eng(torch.zeros(1, 1, size[0], size[1]), size, up)
            if out is not None:
                res[f'out_{sum(emitted)}'] = out.numpy()
            emitted.append(out is not None)
        tail = eng.end(up)
        n_call = sum(emitted)
        for j, out in enumerate(tail):
            res[f'out_{n_call + j}'] = out.numpy()
        save(name, params=json.dumps(dict(thing_list=[1], label_divisor=20000, stuff_area=32,
                                          void_label=0, nms_threshold=0.1, nms_kernel=3,
                                          confidence_thr=0.3, median_kernel_size=ks,
                                          padding_factor=16, coarse_boundaries=(coarse == 4),
                                          upsampling=up, size=list(size), n=D,
                                          emitted=emitted, n_tail=len(tail))), **res)


# ---------------------------------------------------------------------------------------------
# RLE
# ---------------------------------------------------------------------------------------------
def flatten_rle(rle_seg):
    inst, starts, runs = [], [], []
    for cls, attrs in rle_seg.items():
        for lab, a in attrs.items():
            inst.append([cls, lab, *a['box'], len(a['starts'])])
            starts.append(np.asarray(a['starts'], np.int64))
            runs.append(np.asarray(a['runs'], np.int64))
    inst = np.asarray(inst, np.int64).reshape(-1, 7)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
    return inst, cat(starts), cat(runs)


def rle_case(name, pan, labels, L, things, fc):
    pan = np.ascontiguousarray(pan)
    seg = rrle.pan_seg_to_rle_seg(pan, labels, L, things, fc)
    inst, starts, runs = flatten_rle(seg)
    back = rrle.rle_seg_to_pan_seg(seg, pan.shape)
    save(name, in_pan=pan.astype(np.int64), out_inst=inst, out_starts=starts, out_runs=runs,
         out_back=back,
         params=json.dumps(dict(labels=labels, label_divisor=L, thing_list=things,
                                force_connected=fc)))
    print(f'  {name}: {inst.shape[0]} instances, {starts.size} runs')


def matcher_target():
    seg = np.zeros((200, 200), dtype=np.uint32)     # tests/test_matcher.py:6-19 known-answer map
    seg[:16, :16] = 1001
    seg[30:50, 30:50] = 1002
    seg[:10, -10:] = 1003
    seg[-50:, :50] = 1004
    seg[-30:, -30:] = 1005
    seg[100:130, 90:110] = 1006
    return seg


def rle_cases():
    rle_case('rle_matcher_target', matcher_target(), [1], 1000, [1], False)
    rle_case('rle_matcher_target_fc', matcher_target(), [1], 1000, [1], True)

    # SURVEY A5 example: separated blocks of one label, an 8-connected chain, row wrap
    pan = np.zeros((8, 10), np.int64)
    pan[0:2, 0:2] = 1001
    pan[0:2, 5:7] = 1001
    pan[4, 3:8] = 1002
    pan[5, 3:8] = 1002
    pan[6, 8] = 1002
    pan[7, 9] = 1002
    rle_case('rle_a5_fc', pan, [1], 1000, [1], True)
    rle_case('rle_a5_nofc', pan, [1], 1000, [1], False)

    # runs wrapping row ends + full rows + stuff class + two thing classes
    rng = np.random.default_rng(50)
    pan = np.zeros((40, 33), np.int64)
    pan[3:6, :] = 2000                      # stuff class 2 spanning full rows -> one long run
    pan[10, 20:] = 1001
    pan[11, :7] = 1001                      # wraps from col W-1 into col 0 (not 8-connected!)
    pan[20:24, 30:] = 3002
    pan[21:25, :3] = 3002
    pan[30:35, 5:25] = 1007
    pan[31:34, 10:20] = 3001                # hole filled by another class
    pan[38, 1::2] = 1009                    # many single-pixel runs
    pan[39, 0::2] = 1009                    # diagonal 8-connectivity
    rle_case('rle_wrap_fc', pan, [1, 2, 3], 1000, [1, 3], True)
    rle_case('rle_wrap_nofc', pan, [1, 2, 3], 1000, [1, 3], False)
    rle_case('rle_wrap_subset', pan, [3, 1], 1000, [3], True)

    # random blobs: panoptic output of the reference on a synthetic tile, two classes
    d = synth_tile(96, 120, 40, seed=51, semi_axes=(4, 12), sigma=2.5, thing_classes=(1, 2),
                   stuff_classes=(3,))
    pan, _ = rpp.get_panoptic_segmentation(t(d['sem']), t(d['ctr_hmp']), t(d['offsets']), [1, 2],
                                           1000, 16, 0, 0.1, 5)
    pan = pan.numpy()[0, 0]
    rle_case('rle_synth_fc', pan, [1, 2, 3], 1000, [1, 2], True)
    rle_case('rle_synth_nofc', pan, [1, 2, 3], 1000, [1, 2], False)

    # salt-and-pepper: worst case for CCL / run counts
    pan = np.where(rng.random((50, 64)) < 0.5, 1001, 0).astype(np.int64)
    pan[rng.random((50, 64)) < 0.2] = 1002
    rle_case('rle_noise_fc', pan, [1], 1000, [1], True)
    # spiral / U shapes that need many union-find merges
    pan = np.zeros((41, 41), np.int64)
    for r in range(0, 20, 2):
        pan[r, r:41 - r] = 1001
        pan[40 - r, r:41 - r] = 1001
        pan[r:41 - r, r] = 1001
        pan[r + 2:41 - r, 40 - r] = 1001
    rle_case('rle_spiral_fc', pan, [1], 1000, [1], True)
    rle_case('rle_empty', np.zeros((16, 16), np.int64), [1, 2], 1000, [1], True)


# ---------------------------------------------------------------------------------------------
# cross-slice matcher: the reference's rle_matcher / RLEMatcher, forward + backward
# ---------------------------------------------------------------------------------------------
def blob_stack(D, H, W, n_blobs, seed, L=1000):
    """(D,H,W) int64 panoptic-style label maps: ellipsoidal blobs drifting in z, labelled per slice
    1..n by raster order (as the post-processing would), so labels do NOT agree across slices."""
    rng = np.random.default_rng(seed)
    zc, zr = rng.uniform(0, D, n_blobs), rng.uniform(1.5, D / 2, n_blobs)
    cy, cx = rng.uniform(0, H, n_blobs), rng.uniform(0, W, n_blobs)
    vy, vx = rng.normal(0, 1.5, n_blobs), rng.normal(0, 1.5, n_blobs)
    a, b = rng.uniform(4, 14, n_blobs), rng.uniform(4, 14, n_blobs)
    yy, xx = np.mgrid[0:H, 0:W]
    out = np.zeros((D, H, W), np.int64)
    for z in range(D):
        ins = np.zeros((H, W), np.int64)
        for i in np.flatnonzero(np.abs(z - zc) < zr):
            f = np.sqrt(max(0.05, 1 - ((z - zc[i]) / zr[i]) ** 2))
            m = ((yy - cy[i] - vy[i] * (z - zc[i])) / (a[i] * f)) ** 2 + ((xx - cx[i] - vx[i] * (z - zc[i])) / (b[i] * f)) ** 2 <= 1
            ins[m] = i + 1
        out[z] = np.where(ins > 0, L + ins, 0)
    return out


def matcher_cases():
    from empanada.inference import matcher as rmatch  # reference, via shim

    # 1. the reference's own known-answer test (tests/test_matcher.py:6-66)
    target = matcher_target()
    match = np.zeros((200, 200), dtype=np.uint32)
    match[:16, :16] = 1009; match[30:50, 30:50] = 1008; match[:10, -10:] = 1007; match[-50:, :50] = 1006
    match[-30:, 45:80] = 1005; match[-20:, -20:] = 1004; match[100:115, 90:110] = 1003
    match[115:130, 90:110] = 1002; match[50:75, 125:160] = 1001
    out = np.zeros((200, 200), dtype=np.uint32)
    out[:16, :16] = 1001; out[30:50, 30:50] = 1002; out[:10, -10:] = 1003; out[-50:, :50] = 1004
    out[-30:, 45:80] = 1008; out[-20:, -20:] = 1005; out[100:115, 90:110] = 1006
    out[115:130, 90:110] = 1006; out[50:75, 125:160] = 1007
    m = rmatch.RLEMatcher(1, 1000, 0.25, 0.25, True)
    trle = rrle.pan_seg_to_rle_seg(target, [1], 1000, [1], False)
    mrle = rrle.pan_seg_to_rle_seg(match, [1], 1000, [1], False)
    m.initialize_target(trle[1])
    mrle[1] = m(mrle[1], update_target=False)
    assert np.array_equal(rrle.rle_seg_to_pan_seg(mrle, target.shape), out)       # the reference's own assertion
    (ml, al, mi, iou, ioa) = rmatch.rle_matcher(trle[1], rrle.pan_seg_to_rle_seg(match, [1], 1000, [1], False)[1], 0.25,
                                                return_iou=True, return_ioa=True)
    save('matcher_known_answer', target=target.astype(np.int64), match=match.astype(np.int64), out=out.astype(np.int64),
         matched_t=np.asarray(ml[0], np.int64), matched_m=np.asarray(ml[1], np.int64), matched_ious=np.asarray(mi, np.float64),
         iou=np.asarray(iou, np.float64), ioa=np.asarray(ioa, np.float32))

    # 2. forward + backward matching over drifting blob stacks (patterns.py:68-112 recipe)
    for name, (D, H, W, n, seed, fc) in {'matcher_stack_a': (9, 96, 128, 40, 7, True),
                                           'matcher_stack_b': (7, 64, 200, 70, 8, True),
                                           'matcher_stack_nofc': (6, 80, 80, 25, 9, False)}.items():
        vol = blob_stack(D, H, W, n, seed)
        rles = [rrle.pan_seg_to_rle_seg(vol[z], [1], 1000, [1], fc) for z in range(D)]
        res = {'in_vol': vol}
        mt = rmatch.RLEMatcher(1, 1000, 0.25, 0.25, True)
        fwd = []
        for z in range(D):
            seg = {1: rles[z][1]}
            if mt.target_rle is None:
                mt.initialize_target(seg[1])
            else:
                seg[1] = mt(seg[1])
            fwd.append(seg)
            res[f'fwd_{z}'] = rrle.rle_seg_to_pan_seg(seg, (H, W)).astype(np.int64)
            inst, st, ru = flatten_rle(seg)
            res[f'fwd_inst_{z}'], res[f'fwd_starts_{z}'], res[f'fwd_runs_{z}'] = inst, st, ru
        res['fwd_next_label'] = np.int64(mt.next_label)
        mt.target_rle = None
        mt.assign_new = False
        for z in range(D - 1, -1, -1):
            seg = {1: fwd[z][1]}
            if mt.target_rle is None:
                mt.initialize_target(seg[1])
            else:
                seg[1] = mt(seg[1])
            res[f'bwd_{z}'] = rrle.rle_seg_to_pan_seg(seg, (H, W)).astype(np.int64)
            inst, st, ru = flatten_rle(seg)
            res[f'bwd_inst_{z}'], res[f'bwd_starts_{z}'], res[f'bwd_runs_{z}'] = inst, st, ru
        # pairwise matrices of the first two slices, straight from rle_matcher
        (ml, al, mi, iou, ioa) = rmatch.rle_matcher(rles[0][1], rles[1][1], 0.25, return_iou=True, return_ioa=True)
        res.update(pair_matched_t=np.asarray(ml[0], np.int64), pair_matched_m=np.asarray(ml[1], np.int64),
                   pair_ious=np.asarray(mi, np.float64), pair_iou=np.asarray(iou, np.float64), pair_ioa=np.asarray(ioa, np.float32))
        save(name, params=json.dumps(dict(D=D, H=H, W=W, force_connected=fc, label_divisor=1000, class_id=1,
                                          merge_iou_thr=0.25, merge_ioa_thr=0.25)), **res)
        print(f'  {name}: {sum(len(r[1]) for r in rles)} instances in, final labels {len(np.unique(res["bwd_0"])) - 1} in slice 0')


# ---------------------------------------------------------------------------------------------
# tracker: the reference's InstanceTracker along all three axes + its json wire format
# ---------------------------------------------------------------------------------------------
def tracker_cases():
    import tempfile
    from empanada.inference import tracker as rtrack  # reference
    rng = np.random.default_rng(77)
    vol = rng.integers(0, 5, size=(9, 11, 13)).astype(np.int64)
    vol[vol > 0] += 1000
    vol[2:5, 3:9, 4:12] = 1007                          # one big object so runs wrap around row ends
    res = {'in_vol': vol}
    for axis in ('xy', 'xz', 'yz'):
        tr = rtrack.InstanceTracker(1, 1000, vol.shape, axis=axis)
        n = vol.shape[{'xy': 0, 'xz': 1, 'yz': 2}[axis]]
        for i in range(n):
            sl = vol[i] if axis == 'xy' else vol[:, i] if axis == 'xz' else vol[..., i]
            tr.update(rrle.pan_seg_to_rle_seg(np.ascontiguousarray(sl), [1], 1000, [1], False)[1], i)
        tr.finish()
        labs = list(tr.instances.keys())
        res[f'{axis}_labels'] = np.asarray(labs, np.int64)
        res[f'{axis}_boxes'] = np.asarray([tr.instances[l]['box'] for l in labs], np.int64)
        res[f'{axis}_counts'] = np.asarray([len(tr.instances[l]['starts']) for l in labs], np.int64)
        res[f'{axis}_starts'] = np.concatenate([np.asarray(tr.instances[l]['starts'], np.int64) for l in labs])
        res[f'{axis}_runs'] = np.concatenate([np.asarray(tr.instances[l]['runs'], np.int64) for l in labs])
        if axis == 'xy':
            with tempfile.TemporaryDirectory() as d:
                path = os.path.join(d, 't.json')
                tr.write_to_json(path)
                res['xy_json'] = np.frombuffer(open(path, 'rb').read(), dtype=np.uint8)
    save('tracker_axes', **res)


# ---------------------------------------------------------------------------------------------
# orthoplane consensus: the reference's merge_objects_from_trackers / merge_semantic_from_trackers
# ---------------------------------------------------------------------------------------------
def _ball(r):
    g = np.arange(-r, r + 1)
    z, y, x = np.meshgrid(g, g, g, indexing='ij')
    return (z * z + y * y + x * x <= r * r).astype(np.int64)       # skimage.morphology.ball


def _trackers_from_volumes(rtrack, vols, L=1000):
    trs = []
    for vol in vols:
        tr = rtrack.InstanceTracker(1, L, vol.shape, axis='xy')    # tests/test_consensus.py uses 'xy' for all three
        for i, sl in enumerate(vol):
            tr.update(rrle.pan_seg_to_rle_seg(np.ascontiguousarray(sl), [1], L, [1], force_connected=False)[1], i)
        tr.finish()
        trs.append(tr)
    return trs


def _flat_instances(inst):
    labs = list(inst.keys())
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
    return (np.asarray(labs, np.int64), np.asarray([inst[l]['box'] for l in labs], np.int64).reshape(len(labs), -1),
            np.asarray([len(inst[l]['starts']) for l in labs], np.int64),
            cat([np.asarray(inst[l]['starts'], np.int64) for l in labs]), cat([np.asarray(inst[l]['runs'], np.int64) for l in labs]))


def consensus_cases():
    import empanada.array_utils as rau
    from empanada import consensus as rcons
    from empanada.inference import tracker as rtrack
    try:                                            # numba 0.65 cannot compile rle_voting (internal AssertionError):
        rau.rle_voting(np.array([[0, 5], [2, 8]]), 2)
    except Exception:                               # run the reference's own Python body of it instead
        rau.rle_voting = rau.rle_voting.py_func
        rcons.rle_voting = rau.rle_voting
    SETTINGS = [(2, 0.75, False), (2, 0.5, False), (1, 0.75, False), (1, 0.75, True), (3, 0.75, False)]

    def run(name, vols):
        res = {f'in_vol_{i}': v.astype(np.int32) for i, v in enumerate(vols)}
        for k, (vote, thr, bypass) in enumerate(SETTINGS):
            trs = _trackers_from_volumes(rtrack, vols)
            inst = rcons.merge_objects_from_trackers(trs, pixel_vote_thr=vote, cluster_iou_thr=thr, bypass=bypass)
            for key, arr in zip(('labels', 'boxes', 'counts', 'starts', 'runs'), _flat_instances(inst)):
                res[f'obj{k}_{key}'] = arr
            print(f'  {name} objects vote={vote} thr={thr} bypass={bypass}: {len(inst)} instances')
        for k, vote in enumerate((2, 1, 3)):
            trs = _trackers_from_volumes(rtrack, vols)
            for tr in trs:
                if tr.instances:
                    tr.instances = {1001: rcons.merge_instances(tr.instances)}
            inst = rcons.merge_semantic_from_trackers(trs, pixel_vote_thr=vote)
            for key, arr in zip(('labels', 'boxes', 'counts', 'starts', 'runs'), _flat_instances(inst)):
                res[f'sem{k}_{key}'] = arr
        save(name, params=json.dumps(dict(settings=SETTINGS, sem_votes=[2, 1, 3])), **res)

    # 1. the reference's known-answer construction (tests/test_consensus.py:10-60) at half scale
    shape = (50, 50, 50)
    s2 = _ball(10)
    s4 = s2.copy()
    s4[:, 10:, 10:] = 0
    xy, xz, yz = (np.zeros(shape, np.int64) for _ in range(3))
    xy[:21, :21, :21][s2 > 0] = 1001
    xy[8:29, 8:29, 8:29][s2 > 0] = 1002
    xz[:21, :21, :21][s2 > 0] = 1005
    xz[8:29, 8:29, 8:29][s4 > 0] = 1004
    xz[:21, 29:50, 29:50][s2 > 0] = 1006
    yz[:21, :21, :21][s2 > 0] = 1003
    yz[8:29, 8:29, 8:29][s4 > 0] = 1003
    run('consensus_spheres', [xy, xz, yz])

    # 2. many objects seen three times with independent errors: eroded / shifted / split / missing copies
    rng = np.random.default_rng(91)
    shape = (40, 56, 64)
    zz, yy, xx = np.mgrid[0:shape[0], 0:shape[1], 0:shape[2]]
    vols = [np.zeros(shape, np.int64) for _ in range(3)]
    for i in range(26):
        c = np.array([rng.uniform(4, shape[0] - 4), rng.uniform(5, shape[1] - 5), rng.uniform(5, shape[2] - 5)])
        r = rng.uniform(3, 7, 3)
        for v, vol in enumerate(vols):
            if rng.random() < 0.12:
                continue                                            # this view missed the object
            cc = c + rng.normal(0, 0.7, 3)
            rr = r * rng.uniform(0.85, 1.1)
            m = ((zz - cc[0]) / rr[0]) ** 2 + ((yy - cc[1]) / rr[1]) ** 2 + ((xx - cc[2]) / rr[2]) ** 2 <= 1
            lab = 1001 + i + 100 * v
            if rng.random() < 0.15:                                 # a false split: two labels for one object
                vol[m & (xx < cc[2])] = lab
                vol[m & (xx >= cc[2])] = lab + 50
            else:
                vol[m] = lab
    run('consensus_blobs', vols)


# ---------------------------------------------------------------------------------------------
# orthoplane chain (BASELINE configs[3] in miniature): the reference's scripts/pdl_inference3d.py flow
# — Render engine per slice along xy / xz / yz, RLE, forward + backward matching, trackers, filters,
# instance consensus, filling — on replayed head tensors of one small blob volume
# ---------------------------------------------------------------------------------------------
def _patch_rle_voting():
    import empanada.array_utils as rau
    from empanada import consensus as rcons
    try:
        rau.rle_voting(np.array([[0, 5], [2, 8]]), 2)
    except Exception:
        rau.rle_voting = rau.rle_voting.py_func
        rcons.rle_voting = rau.rle_voting


def ortho_cases():
    from empanada.inference import patterns as rpat, filters as rfilt
    from empanada_b200.synth import synth_blob_volume, synth_heads
    _patch_rle_voting()
    shape, L, seed = (36, 40, 44), 1000, 7
    P = dict(thing_list=[1], labels=[1], label_divisor=L, stuff_area=16, void_label=0, nms_threshold=0.1, nms_kernel=3,
             confidence_thr=0.3, median_kernel_size=3, padding_factor=16, coarse_boundaries=False,
             merge_iou_thr=0.25, merge_ioa_thr=0.25, min_size=40, min_span=3, pixel_vote_thr=2,
             cluster_iou_thr=0.75, shape=list(shape))
    vol = synth_blob_volume(shape, 12, seed)
    axes = {'xy': 0, 'xz': 1, 'yz': 2}
    res = {'in_vol': vol}
    trackers = rpat.create_axis_trackers(axes, P['labels'], L, shape)
    for name, axis in axes.items():
        n = shape[axis]
        h, w = [d for a, d in enumerate(shape) if a != axis]
        Hp, Wp = -(-h // 16) * 16, -(-w // 16) * 16
        outs = []
        for i in range(n):
            padded = np.zeros((Hp, Wp), np.int32)
            padded[:h, :w] = np.take(vol, i, axis=axis)
            hd = synth_heads(padded, np.random.default_rng([seed, axis, i]))
            o = {'sem_logits': logit(hd['sem_prob']), 'ctr_hmp': hd['ctr_hmp'], 'offsets': hd['offsets']}
            for k, v in o.items():
                res[f'in_{name}_{i}_{k}'] = v
            outs.append({k: t(v) for k, v in o.items()})
        eng = reng.PanopticDeepLabRenderEngine3d(
            ReplayModel(outs), thing_list=P['thing_list'], median_kernel_size=P['median_kernel_size'],
            label_divisor=L, stuff_area=P['stuff_area'], void_label=P['void_label'], nms_threshold=P['nms_threshold'],
            nms_kernel=P['nms_kernel'], confidence_thr=P['confidence_thr'], padding_factor=P['padding_factor'],
            coarse_boundaries=False)
        matchers = rpat.create_matchers(P['thing_list'], L, P['merge_iou_thr'], P['merge_ioa_thr'])
        rle_stack = []

        def take(pan):
            seg = rrle.pan_seg_to_rle_seg(pan.squeeze().cpu().numpy(), P['labels'], L, P['thing_list'], force_connected=True)
            rle_stack.append(rpat.apply_matchers(seg, matchers))

        for i in range(n):
            pan = eng(torch.zeros(1, 1, h, w), (h, w), upsampling=1)
            if pan is not None:
                take(pan)
        for pan in eng.end(1):
            take(pan)
        assert len(rle_stack) == n
        for index, seg in rpat.backward_matching(rle_stack, matchers, n):
            rpat.update_trackers(seg, index, trackers[name])
        # the script fills the per-axis panoptic stack BEFORE finish_tracking (pdl_inference3d.py:198-208), which
        # raises in numpy_fill_instances (starts are still lists of arrays); the working order is used here
        rpat.finish_tracking(trackers[name])
        stack = np.zeros(shape, np.uint32)
        rpat.fill_panoptic_volume(stack, trackers[name])
        res[f'out_{name}_stack'] = stack
        for tr in trackers[name]:
            rfilt.remove_small_objects(tr, min_size=P['min_size'])
            rfilt.remove_pancakes(tr, min_span=P['min_span'])
            for key, arr in zip(('labels', 'boxes', 'counts', 'starts', 'runs'), _flat_instances(tr.instances)):
                res[f'out_{name}_{key}'] = arr
        print(f'  ortho {name}: {len(trackers[name][0].instances)} instances after filters, '
              f'{len(np.unique(stack)) - 1} labels painted')
    cons = rpat.create_instance_consensus(rpat.get_axis_trackers_by_class(trackers, 1), P['pixel_vote_thr'],
                                          P['cluster_iou_thr'], False)
    rfilt.remove_small_objects(cons, min_size=P['min_size'])
    rfilt.remove_pancakes(cons, min_span=P['min_span'])
    for key, arr in zip(('labels', 'boxes', 'counts', 'starts', 'runs'), _flat_instances(cons.instances)):
        res[f'out_cons_{key}'] = arr
    cvol = np.zeros(shape, np.uint32)
    rpat.fill_volume(cvol, cons.instances)
    res['out_cons_vol'] = cvol
    print(f'  ortho consensus: {len(cons.instances)} instances ({len(np.unique(vol)) - 1} blobs in the volume)')
    save('ortho_chain', params=json.dumps(P), **res)


def stack_rle_cases():
    """The stack path end to end, by the reference alone: PanopticDeepLabRenderEngine3d over a replayed stack (recursive
    median queue, coarse instance cells, merge, crop: engines.py:327-394), then pan_seg_to_rle_seg(force_connected) per
    emitted slice (patterns.py:93-95) — inputs and every slice's RLE tables."""
    for name, ks, up, D, size, things, ncls in (('stack_rle_ks3', 3, 1, 9, (60, 75), [1], 1),
                                                ('stack_rle_ks5_up2', 5, 2, 8, (118, 150), [1], 1),
                                                ('stack_rle_ks1_mc', 1, 1, 4, (64, 80), [1, 2], 3)):
        H, W = 64, 80
        base = stack_inputs(D, H * up, W * up, seed=52, coarse=1, n_classes=ncls)
        ins = []
        for z in range(D):
            o = dict(base[z])
            dd = synth_tile(H // 4, W // 4, 10, seed=53 + z // 3, semi_axes=(2, 4), sigma=1.2)
            o['ctr_hmp'] = dd['ctr_hmp'].astype(np.float32)
            o['offsets'] = (dd['offsets'] * 4).astype(np.float32)
            ins.append(o)
        model = ReplayModel([{k: t(v) for k, v in o.items()} for o in ins])
        eng = reng.PanopticDeepLabRenderEngine3d(model, things, 20000, 32, 0, 0.1, 3, 0.3, ks, 16, coarse_boundaries=True)
        pans = []
        for z in range(D):
            out = eng(torch.zeros(1, 1, size[0], size[1]), size, up)
            if out is not None:
                pans.append(out)
        pans += eng.end(up)
        assert len(pans) == D
        labels = things + ([3] if ncls == 3 else [])
        res = {}
        for z, o in enumerate(ins):
            for k, v in o.items():
                res[f'in_{z}_{k}'] = v
            seg = rrle.pan_seg_to_rle_seg(pans[z].squeeze().numpy(), labels, 20000, things, True)
            inst, starts, runs = flatten_rle(seg)
            res[f'out_{z}_inst'], res[f'out_{z}_starts'], res[f'out_{z}_runs'] = inst, starts, runs
        n_inst = sum(res[f'out_{z}_inst'].shape[0] for z in range(D))
        save(name, params=json.dumps(dict(thing_list=things, labels=labels, label_divisor=20000, stuff_area=32, void_label=0,
                                          nms_threshold=0.1, nms_kernel=3, confidence_thr=0.3, median_kernel_size=ks,
                                          padding_factor=16, upsampling=up, size=list(size), n=D)), **res)
        print(f'  {name}: {D} slices, {n_inst} instances')


if __name__ == '__main__':
    which = sys.argv[1:] or ['pp', 'merge', 'engine', 'rle', 'matcher', 'tracker', 'consensus', 'ortho', 'stack']
    if 'stack' in which:
        stack_rle_cases()
    if 'ortho' in which:
        ortho_cases()
    if 'pp' in which:
        pp_cases()
    if 'merge' in which:
        merge_cases()
    if 'engine' in which:
        engine_cases()
    if 'rle' in which:
        rle_cases()
    if 'matcher' in which:
        matcher_cases()
    if 'tracker' in which:
        tracker_cases()
    if 'consensus' in which:
        consensus_cases()
