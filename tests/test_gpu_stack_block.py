"""The batched stack path (csrc/stack_block.cu): emp_median_chain / emp_median_chain_repair against the plane-by-plane
restatement of the reference's recursive queue (inference/stack.median_chain with torch.median, which
tests/test_stack_host.py pins to the reference's _MedianQueue fixtures), and emp_stack_block — RLE tables cut from
the code map — against the golden-pinned per-slice path (Render engine -> emp_rle) and the oracle."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tmedian(window):
    return torch.median(torch.cat(window, dim=0), dim=0, keepdim=True).values


def _harden(p, thr):
    """engines.py:114-121 in torch: C > 1 first arg-max (NaN is the maximum), C == 1 p >= thr"""
    if p.size(1) > 1:
        return torch.argmax(p, dim=1).to(torch.uint8)[0]
    return (p >= thr).to(torch.uint8)[0, 0]


def _run_chain(dev, raw, z0, z1, depth, ks, carry, thr, want_carry=True):
    from empanada_b200 import _cabi as C
    mid = ks // 2
    planes = [raw[z] for z in range(z0, min(depth, z1 + mid))]
    Cn, H, W = planes[0].shape[1:]
    hw = H * W
    n = z1 - z0
    out = [torch.full((Cn * hw,), -7.0, dtype=torch.float32, device=dev) for _ in range(mid)]
    words = [p.data_ptr() for p in planes] + [t.data_ptr() for t in out] + [t.data_ptr() for t in (carry or [])]
    tab = torch.tensor(words, dtype=torch.int64).to(dev)
    sem8 = torch.full((n, hw), 99, dtype=torch.uint8, device=dev)
    best = torch.empty((n, hw), dtype=torch.float32, device=dev) if Cn > 1 else None
    vp = ctypes.c_void_p
    with torch.cuda.device(dev):
        C.check(C.lib().emp_median_chain(vp(tab.data_ptr()), n, len(planes), z0, depth, ks, Cn, hw,
                                         vp(tab.data_ptr() + 8 * (len(planes) + mid)) if carry else None, float(thr),
                                         vp(sem8.data_ptr()), hw, vp(best.data_ptr()) if best is not None else None,
                                         vp(tab.data_ptr() + 8 * len(planes)) if (want_carry and mid) else None, None,
                                         C.stream_ptr(dev)))
    torch.cuda.synchronize(dev)
    return sem8.view(n, H, W), [t.view(1, Cn, H, W) for t in out]


@pytest.mark.parametrize('ks,Cn,shape', [(1, 1, (40, 64)), (3, 1, (40, 64)), (3, 1, (37, 51)), (5, 1, (24, 36)),
                                         (7, 1, (16, 20)), (3, 3, (20, 28)), (5, 2, (9, 13)), (15, 1, (8, 12))])
@pytest.mark.parametrize('z0,z1,depth', [(0, 17, 17), (0, 9, 30), (8, 19, 30), (21, 30, 30)])
def test_median_chain_block(ks, Cn, shape, z0, z1, depth, cuda_device):
    from empanada_b200.inference import stack
    if depth < ks:
        pytest.skip('stack shallower than the kernel')
    mid = ks // 2
    rng = np.random.default_rng(ks * 1000 + z0 * 10 + Cn)
    H, W = shape
    raw = {}
    for z in range(max(z0 - mid, 0), min(depth, z1 + mid)):
        a = np.round(rng.random((1, Cn, H, W), dtype=np.float32) * 32) / 32          # many ties
        if z % 5 == 2:
            a[0, 0, 3, 5] = np.nan                                                  # torch.median propagates NaN
        raw[z] = torch.from_numpy(a).to(cuda_device)
    thr = 0.4
    carry = [torch.from_numpy(rng.random((1, Cn, H, W), dtype=np.float32)).to(cuda_device) for _ in range(mid)] if z0 > 0 else []
    want, want_carry = stack.median_chain(raw, z0, z1, depth, ks, carry, _tmedian)
    sem8, got_carry = _run_chain(cuda_device, raw, z0, z1, depth, ks, [c.reshape(-1) for c in carry], thr)
    for i, z in enumerate(range(z0, z1)):
        assert torch.equal(sem8[i], _harden(want[z], thr)), f'slice {z}'
    for a, b in zip(got_carry, want_carry):
        assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))


@pytest.mark.parametrize('ks', [3, 5, 9])
@pytest.mark.parametrize('kind', ['noise', 'smooth', 'constant', 'nan'])   # 'constant': a filter that never forgets
def test_median_chain_repair(ks, kind, cuda_device):
    """chain from a guessed carry + repair from the true one == chain from the true carry (class maps and outgoing
    carry), and the changed flag says whether the outgoing carry moved."""
    from empanada_b200 import _cabi as C
    from empanada_b200.inference import stack
    dev = cuda_device
    mid = ks // 2
    rng = np.random.default_rng(ks)
    H, W, depth, z0, z1 = 33, 48, 40, 11, 29
    raw = {}
    base = rng.random((1, 1, H, W), dtype=np.float32)
    for z in range(z0, z1 + mid):
        if kind == 'noise':
            a = rng.random((1, 1, H, W), dtype=np.float32)
        elif kind == 'smooth':
            a = base + 0.01 * z + 0.02 * rng.random((1, 1, H, W), dtype=np.float32)
        elif kind == 'constant':                                # alternating 0 / 10: every window's middle values are the carried ones,
            a = np.full((1, 1, H, W), 10.0 * (z % 2), np.float32)   # so the filter never forgets its start
        else:
            a = rng.random((1, 1, H, W), dtype=np.float32)
            a[0, 0, 2, 2] = np.nan
        raw[z] = torch.from_numpy(a.astype(np.float32)).to(dev)
    true_carry = [torch.from_numpy(rng.random((1, 1, H, W), dtype=np.float32) * (3.0 if kind == 'constant' else 1.0)).to(dev) for _ in range(mid)]
    thr = 0.5
    want_sem, want_carry = _run_chain(dev, raw, z0, z1, depth, ks, [c.reshape(-1) for c in true_carry], thr)
    guess = [raw[z0].reshape(-1)] * mid
    sem8, carry_out = _run_chain(dev, raw, z0, z1, depth, ks, guess, thr)
    spec_carry = [c.clone() for c in carry_out]
    planes = [raw[z] for z in range(z0, z1 + mid)]
    n, hw = z1 - z0, H * W
    words = [p.data_ptr() for p in planes] + [g.data_ptr() for g in guess] + [c.data_ptr() for c in true_carry] + \
            [c.data_ptr() for c in carry_out]
    tab = torch.tensor(words, dtype=torch.int64).to(dev)
    changed = torch.zeros((1,), dtype=torch.int32, device=dev)
    sem8 = sem8.contiguous().view(n, hw)
    vp = ctypes.c_void_p
    P = len(planes)
    with torch.cuda.device(dev):
        C.check(C.lib().emp_median_chain_repair(vp(tab.data_ptr()), n, P, z0, depth, ks, hw, vp(tab.data_ptr() + 8 * P),
                                                vp(tab.data_ptr() + 8 * (P + mid)), thr, vp(sem8.data_ptr()), hw,
                                                vp(tab.data_ptr() + 8 * (P + 2 * mid)), vp(changed.data_ptr()), None, C.stream_ptr(dev)))
    torch.cuda.synchronize(dev)
    assert torch.equal(sem8.view(n, H, W), want_sem)
    moved = any(not torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)) for a, b in zip(spec_carry, want_carry))
    for a, b in zip(carry_out, want_carry):
        assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))
    assert bool(changed.item()) == moved
    if kind == 'constant':
        assert moved                                            # the case the settle rounds exist for
    if kind == 'smooth':
        assert not moved


def _rle_equal(a, b, ctx=''):
    assert list(a.keys()) == list(b.keys()), ctx
    for c in a:
        assert list(a[c].keys()) == list(b[c].keys()), f'{ctx} class {c}'
        for lab in a[c]:
            assert tuple(a[c][lab]['box']) == tuple(b[c][lab]['box']), f'{ctx} {lab}'
            np.testing.assert_array_equal(a[c][lab]['starts'], b[c][lab]['starts'])
            np.testing.assert_array_equal(a[c][lab]['runs'], b[c][lab]['runs'])


def _per_slice_reference(e, probs, heads, sizes, ks, depth, labels, force_connected, upsampling=1):
    """the golden-pinned path: chain plane by plane (emp_median_harden), fused Render post-process, crop, emp_rle"""
    from empanada_b200.inference import engines as eng, rle, stack
    raw = {z: probs[z] for z in range(depth)}
    filt, _ = stack.median_chain(raw, 0, depth, depth, ks, [], lambda w: eng.median_harden(w, 0.0)[0])
    out = []
    for z in range(depth):
        pan = e._fused_postprocess(filt[z], heads[z]['ctr_hmp'], heads[z]['offsets'], upsampling)
        pan = pan[..., :sizes[0], :sizes[1]]
        out.append(rle.pan_seg_to_rle_seg(pan, labels, e.label_divisor, e.thing_list, force_connected))
    return out


@pytest.mark.parametrize('streamed', [False, True, 'side'])
@pytest.mark.parametrize('case', ['plain', 'crop', 'odd', 'multiclass', 'noccl', 'void7', 'block1', 'up2'])
def test_stack_block_matches_per_slice_path(case, streamed, cuda_device):
    from empanada_b200.inference import engines as eng, stack
    from empanada_b200.synth import synth_stack_slices
    dev = cuda_device
    D, H, W, ks, up = 9, 192, 256, 3, 1
    size = (H, W)
    kw = dict(thing_list=[1], label_divisor=20000, stuff_area=64, void_label=0, nms_threshold=0.1, nms_kernel=3,
              confidence_thr=0.3, coarse_boundaries=True)
    labels, fc, block = [1], True, 4
    if case == 'crop':
        size = (H - 21, W - 70)
    elif case == 'odd':
        H, W = 172, 236                                       # W % 8 != 0: scalar code loads; coarse 43 x 59
        size = (H - 3, W - 1)
    elif case == 'noccl':
        fc = False
    elif case == 'void7':
        kw.update(void_label=7, label_divisor=1000)
        labels = [0, 1]                                       # void pixels (7) fall into class 0's label range
    elif case == 'block1':
        block = 1
    elif case == 'up2':
        up = 2
    sl = list(synth_stack_slices(D, H // up, W // up, 40, seed=5, coarse=4, sigma=4.0 / up, z_extent=(4, 12), semi_axes=(6, 20)))
    heads = [{k: torch.from_numpy(s[k]).to(dev) for k in ('sem_prob', 'ctr_hmp', 'offsets')} for s in sl]
    probs = [h['sem_prob'] for h in heads]
    if up == 2:                                               # full-res probabilities at twice the coarse maps' scale
        probs = [torch.nn.functional.interpolate(p, scale_factor=2, mode='bilinear') for p in probs]
    if case == 'multiclass':
        # three channels: background, thing class 1, stuff class 2 (a band across the slice), thing class 3
        kw.update(thing_list=[1, 3], stuff_area=300)
        labels = [1, 2, 3]
        new = []
        for z, p in enumerate(probs):
            q = torch.zeros((1, 4, H, W), device=dev)
            q[:, 1] = p[:, 0] * (torch.arange(W, device=dev)[None, None] < W // 2)
            q[:, 3] = p[:, 0] * (torch.arange(W, device=dev)[None, None] >= W // 2)
            q[:, 2, 10 + z:30 + z] = 0.5
            q[:, 0] = 0.3
            new.append(q.contiguous())
        probs = new
    e = eng.PanopticDeepLabRenderEngine(torch.nn.Identity(), **kw)
    want = _per_slice_reference(e, probs, heads, size, ks, D, labels, fc, up)
    shard = stack.StackShard(e, labels=labels, depth=D, median_kernel_size=ks, upsampling=up, force_connected=fc, block=block,
                             stream=torch.cuda.Stream(dev) if streamed == 'side' else None)
    early = 0
    for z in shard.slices():
        # streamed: sub-blocks are chained and cut into tables as soon as their slices (+ look-ahead) are in, and their heads
        # forgotten (fresh tensors per slice, so that a premature release would show)
        if streamed:
            shard.add(z, probs[z].clone(), heads[z]['ctr_hmp'].clone(), heads[z]['offsets'].clone(), size=size)
            early += shard.advance()
        else:
            shard.add(z, probs[z], heads[z]['ctr_hmp'], heads[z]['offsets'], size=size)
    if streamed:
        assert early >= (D - 1) // block - 1 and 'sem' not in shard.heads[0]
    got = shard.finish()
    assert sorted(got.keys()) == list(range(D))
    assert sum(len(v) for w in want for v in w.values()) > 20
    for z in range(D):
        _rle_equal(got[z], want[z], f'{case} slice {z}')
    n_inst, n_runs = got.counts()
    assert n_inst == sum(len(v) for w in want for v in w.values())
    assert n_runs == sum(len(a['starts']) for w in want for v in w.values() for a in v.values())
    # the row-run tables kept for match() / fill(): painting them reproduces the label maps
    from empanada_b200.inference import rle
    vol = shard.fill(torch.int64).cpu().numpy()
    for z in range(D):
        np.testing.assert_array_equal(vol[z], rle.rle_seg_to_pan_seg(want[z], size).astype(np.int64))


@pytest.mark.parametrize('caps', [(40, 4096), (16384, 3)])
def test_stack_block_table_overflow_falls_back_per_slice(caps, cuda_device):
    """A slice with more row-runs / instances than the deferred tables hold is flagged by the kernels and redone
    synchronously through the per-slice path (tables grown to fit); the other slices of the block keep their batched
    results.  Everything must still equal the per-slice reference path."""
    from empanada_b200.inference import engines as eng, stack
    from empanada_b200.synth import synth_stack_slices
    dev = cuda_device
    D, H, W = 6, 128, 192
    sl = list(synth_stack_slices(D, H, W, 30, seed=15, coarse=4, sigma=4.0, z_extent=(3, 9), semi_axes=(5, 16)))
    heads = [{k: torch.from_numpy(s[k]).to(dev) for k in ('sem_prob', 'ctr_hmp', 'offsets')} for s in sl]
    heads[2]['sem_prob'] = torch.full_like(heads[2]['sem_prob'], 0.1)         # one slice with nothing in it stays under any capacity
    probs = [h['sem_prob'] for h in heads]
    kw = dict(thing_list=[1], label_divisor=20000, stuff_area=64, void_label=0, nms_threshold=0.1, nms_kernel=3,
              confidence_thr=0.3, coarse_boundaries=True)
    e = eng.PanopticDeepLabRenderEngine(torch.nn.Identity(), **kw)
    want = _per_slice_reference(e, probs, heads, (H, W), 1, D, [1], True)
    shard = stack.StackShard(e, labels=[1], depth=D, median_kernel_size=1, run_cap=caps[0], inst_cap=caps[1])
    for z in shard.slices():
        shard.add(z, probs[z], heads[z]['ctr_hmp'], heads[z]['offsets'], size=(H, W))
    got = shard.finish()
    assert shard.tables_['bad'].any() and not shard.tables_['bad'].all()
    for z in range(D):
        _rle_equal(got[z], want[z], f'slice {z}')
    assert got.counts()[0] == sum(len(v) for w in want for v in w.values())


def test_stack_block_vs_oracle_row_wrap(cuda_device):
    """Runs that reach the last column continue in column 0 of the next row (array_utils.rle_encode): full-width
    things, checked against the oracle's restatement of rle.py on the oracle's own panoptic maps."""
    import oracle
    from empanada_b200.inference import engines as eng, stack
    dev = cuda_device
    D, H, W = 3, 64, 128
    rng = np.random.default_rng(3)
    heads, probs = [], []
    for z in range(D):
        p = np.full((1, 1, H, W), 0.1, np.float32)
        p[0, 0, 8:20, :] = 0.9                                  # a full-width band: one run per instance after merging
        p[0, 0, 30:40, 100:] = 0.9                              # touches the right edge only
        p[0, 0, 31:41, :17] = 0.9                               # touches the left edge one row lower: wraps
        hm = np.zeros((1, 1, H // 4, W // 4), np.float32)
        hm[0, 0, 3, 10] = 1.0                                   # one center: every thing pixel takes id 1
        off = rng.normal(0, 0.3, (1, 2, H // 4, W // 4)).astype(np.float32)
        heads.append({'ctr_hmp': torch.from_numpy(hm).to(dev), 'offsets': torch.from_numpy(off).to(dev)})
        probs.append(torch.from_numpy(p).to(dev))
    kw = dict(thing_list=[1], label_divisor=1000, stuff_area=0, void_label=0, nms_threshold=0.1, nms_kernel=3, confidence_thr=0.5)
    e = eng.PanopticDeepLabRenderEngine(torch.nn.Identity(), **kw)
    shard = stack.StackShard(e, labels=[1], depth=D, median_kernel_size=1)
    for z in shard.slices():
        shard.add(z, probs[z], heads[z]['ctr_hmp'], heads[z]['offsets'], size=(H, W))
    got = shard.finish()
    for z in range(D):
        pan = e._fused_postprocess(probs[z], heads[z]['ctr_hmp'], heads[z]['offsets'], 1)[0].cpu().numpy()
        want = oracle.pan_seg_to_rle_seg(pan, [1], 1000, [1], True)
        _rle_equal(got[z], want, f'slice {z}')
        assert any(int(s % W + r) > W for a in got[z][1].values() for s, r in zip(a['starts'], a['runs']))   # a run crossing a row end


@pytest.mark.parametrize('name', ['stack_rle_ks3', 'stack_rle_ks5_up2', 'stack_rle_ks1_mc'])
@pytest.mark.parametrize('block', [128, 3])
def test_stack_shard_vs_reference_stack_of_rle(name, block, cuda_device):
    """StackShard against fixtures the REFERENCE produced for the whole stack path (tests/golden/make_golden.py stack:
    its PanopticDeepLabRenderEngine3d over a replayed stack, then its pan_seg_to_rle_seg(force_connected=True) per
    slice): every slice's labels, boxes, starts and runs."""
    from conftest import load_golden
    from empanada_b200.inference import engines as eng, stack
    g = load_golden(name)
    p = g['params']
    D = p['n']
    e = eng.PanopticDeepLabRenderEngine(torch.nn.Identity(), thing_list=p['thing_list'], label_divisor=p['label_divisor'],
                                        stuff_area=p['stuff_area'], void_label=p['void_label'], nms_threshold=p['nms_threshold'],
                                        nms_kernel=p['nms_kernel'], confidence_thr=p['confidence_thr'], coarse_boundaries=True)
    shard = stack.StackShard(e, labels=p['labels'], depth=D, median_kernel_size=p['median_kernel_size'], upsampling=p['upsampling'],
                             force_connected=True, block=block)
    for z in shard.slices():
        prob = eng.logits_to_prob(torch.from_numpy(g[f'in_{z}_sem_logits']).to(cuda_device))
        shard.add(z, prob, torch.from_numpy(g[f'in_{z}_ctr_hmp']).to(cuda_device), torch.from_numpy(g[f'in_{z}_offsets']).to(cuda_device),
                  size=tuple(p['size']))
    got = shard.finish()
    total = 0
    for z in range(D):
        inst, starts, runs = g[f'out_{z}_inst'], g[f'out_{z}_starts'], g[f'out_{z}_runs']
        want, at = {int(l): {} for l in p['labels']}, 0
        for row in inst:
            n = int(row[6])
            want[int(row[0])][int(row[1])] = {'box': tuple(int(v) for v in row[2:6]), 'starts': starts[at:at + n], 'runs': runs[at:at + n]}
            at += n
        _rle_equal(got[z], want, f'{name} slice {z}')
        total += inst.shape[0]
    assert got.counts()[0] == total


def test_sharded_stack_on_two_gpus_matches_single_block():
    """tests/mp_stack_match.py under torchrun on 2 GPUs (NCCL): every rank's finish + match + fill must equal the single
    block.  Skipped on boxes with one GPU (the bench lines carry the same check at every N)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29541', os.path.join(root, 'tests', 'mp_stack_match.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'mp_stack_match world=2: OK' in r.stdout


def test_stack_block_degenerate_slices(cuda_device):
    """Slices with nothing in them (all background: every strip flagged), with thing pixels but no center (ids 0 -> void),
    and a normal one, in one block."""
    from empanada_b200.inference import engines as eng, stack
    from empanada_b200.synth import synth_stack_slices
    dev = cuda_device
    D, H, W = 4, 128, 192
    sl = list(synth_stack_slices(D, H, W, 25, seed=33, coarse=4, sigma=4.0, z_extent=(3, 9), semi_axes=(5, 16)))
    heads = [{k: torch.from_numpy(s[k]).to(dev) for k in ('sem_prob', 'ctr_hmp', 'offsets')} for s in sl]
    heads[0]['sem_prob'] = torch.full_like(heads[0]['sem_prob'], 0.1)
    heads[1]['ctr_hmp'] = torch.zeros_like(heads[1]['ctr_hmp'])              # things, but no center anywhere
    heads[2]['sem_prob'] = torch.full_like(heads[2]['sem_prob'], 0.9)        # every pixel a thing pixel: no strip is flagged
    probs = [h['sem_prob'] for h in heads]
    kw = dict(thing_list=[1], label_divisor=20000, stuff_area=64, void_label=0, nms_threshold=0.1, nms_kernel=3,
              confidence_thr=0.3, coarse_boundaries=True)
    e = eng.PanopticDeepLabRenderEngine(torch.nn.Identity(), **kw)
    want = _per_slice_reference(e, probs, heads, (H, W), 1, D, [1], True)
    shard = stack.StackShard(e, labels=[1], depth=D, median_kernel_size=1)
    for z in shard.slices():
        shard.add(z, probs[z], heads[z]['ctr_hmp'], heads[z]['offsets'], size=(H, W))
    got = shard.finish()
    assert got[0] == {1: {}} and got[1] == {1: {}}
    for z in range(D):
        _rle_equal(got[z], want[z], f'slice {z}')


def test_stack_block_4096_plane(cuda_device):
    """A 4096 x 4096 plane (coarse 1024 x 1024): 65536-entry run tables, 32-bit flat indices near their upper half."""
    from empanada_b200.inference import engines as eng, stack
    from empanada_b200.synth import synth_stack_slices
    dev = cuda_device
    D, H, W = 2, 4096, 4096
    sl = list(synth_stack_slices(D, H, W, 300, seed=44, coarse=4, z_extent=(2, 6)))
    heads = [{k: torch.from_numpy(s[k]).to(dev) for k in ('sem_prob', 'ctr_hmp', 'offsets')} for s in sl]
    probs = [h['sem_prob'] for h in heads]
    kw = dict(thing_list=[1], label_divisor=20000, stuff_area=64, void_label=0, nms_threshold=0.1, nms_kernel=3,
              confidence_thr=0.3, coarse_boundaries=True)
    e = eng.PanopticDeepLabRenderEngine(torch.nn.Identity(), **kw)
    want = _per_slice_reference(e, probs, heads, (H, W), 1, D, [1], True)
    shard = stack.StackShard(e, labels=[1], depth=D, median_kernel_size=1)
    for z in shard.slices():
        shard.add(z, probs[z], heads[z]['ctr_hmp'], heads[z]['offsets'], size=(H, W))
    got = shard.finish()
    assert sum(len(w[1]) for w in want) > 50
    for z in range(D):
        _rle_equal(got[z], want[z], f'slice {z}')
