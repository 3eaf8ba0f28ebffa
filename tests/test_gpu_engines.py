"""B200: the engine classes (median queue, hardening, fused post-processing, Render-engine coarse
path) replaying the head tensors of reference-generated fixtures through a fake model."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from empanada_b200.inference import engines as eng

pytestmark = pytest.mark.gpu


class ReplayModel(torch.nn.Module):
    def __init__(self, outputs, device):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1, device=device))
        self.outputs = outputs
        self.i = 0

    def forward(self, image, render_steps=None, interpolate_ins=None):
        out = {k: v.clone() for k, v in self.outputs[self.i].items()}
        self.i += 1
        return out


def _inputs(g, n, dev):
    return [{k: torch.from_numpy(g[f'in_{z}_{k}']).to(dev) for k in ('sem_logits', 'ctr_hmp', 'offsets')}
            for z in range(n)]


def _mismatch(got, want):
    got = got.cpu().numpy()
    assert got.shape == want.shape, (got.shape, want.shape)
    return int((got != want).sum())


def test_engine2d(cuda_device):
    g = load_golden('engine2d')
    p = g['params']
    model = ReplayModel(_inputs(g, p['n'], cuda_device), cuda_device)
    e = eng.PanopticDeepLabEngine(model, p['thing_list'], p['label_divisor'], p['stuff_area'], p['void_label'],
                                  p['nms_threshold'], p['nms_kernel'], p['confidence_thr'])
    for z in range(p['n']):
        out = e(torch.zeros(1, 1, 64, 80))
        assert out.dtype == torch.int64
        assert _mismatch(out, g[f'out_{z}']) == 0


def test_config1_tile_matches_reference_run(cuda_device):
    """BASELINE configs[0]: the 1024^2 tile the reference's ResNet-50 PDL + PanopticDeepLabEngine was run on
    (tests/golden/make_config1.py); the CNN's outputs are regenerated from the recorded constants + seed."""
    from empanada_b200.synth import CONFIG1, config1_heads
    g = load_golden('config1')
    heads = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_device) for k, v in config1_heads(g['consts']).items()}

    class Recorded(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.anchor = torch.nn.Parameter(torch.zeros(1, device=cuda_device))       # the engine asks the model for its device

        def forward(self, x):
            return dict(heads)

    e = eng.PanopticDeepLabEngine(Recorded(), **CONFIG1)
    out = e(torch.zeros(1, 1, 1024, 1024, device=cuda_device))
    assert out.dtype == torch.int64 and tuple(out.shape) == (1, 1, 1024, 1024)
    assert _mismatch(out, g['pan'].astype(np.int64)) == 0


@pytest.mark.parametrize('name', golden_names('engine3d_'))
def test_engine3d(name, cuda_device):
    g = load_golden(name)
    p = g['params']
    model = ReplayModel(_inputs(g, p['n'], cuda_device), cuda_device)
    e = eng.PanopticDeepLabEngine3d(model, p['thing_list'], p['label_divisor'], p['stuff_area'], p['void_label'],
                                    p['nms_threshold'], p['nms_kernel'], p['confidence_thr'], p['median_kernel_size'])
    outs, emitted = [], []
    for z in range(p['n']):
        o = e(torch.zeros(1, 1, 48, 64))
        emitted.append(o is not None)
        if o is not None:
            outs.append(o)
    outs += e.end()
    assert emitted == p['emitted'] and len(outs) == sum(p['emitted']) + p['n_tail']
    for i, o in enumerate(outs):
        assert _mismatch(o, g[f'out_{i}']) == 0, f'slice {i}'


@pytest.mark.parametrize('name', golden_names('render3d_'))
def test_render_engine3d(name, cuda_device):
    g = load_golden(name)
    p = g['params']
    size, up = tuple(p['size']), p['upsampling']
    for mode in ('fused', 'reference_methods'):
        model = ReplayModel(_inputs(g, p['n'], cuda_device), cuda_device)
        e = eng.PanopticDeepLabRenderEngine3d(model, p['thing_list'], p['label_divisor'], p['stuff_area'],
                                              p['void_label'], p['nms_threshold'], p['nms_kernel'],
                                              p['confidence_thr'], p['median_kernel_size'], p['padding_factor'],
                                              p['coarse_boundaries'])
        outs = []
        if mode == 'fused':
            for z in range(p['n']):
                o = e(torch.zeros(1, 1, *size), size, up)
                if o is not None:
                    outs.append(o)
            outs += e.end(up)
        else:       # the reference-shaped public methods: get_instance_cells -> postprocess
            from empanada_b200.inference.postprocess import factor_pad
            for z in range(p['n']):
                image = e.to_model_device(factor_pad(torch.zeros(1, 1, *size), e.padding_factor))
                mo = e.infer(image, 2)
                mo['size'] = size
                e.enqueue(mo)
                m = e.get_next(keys=['sem'])
                if m is not None:
                    cells = e.get_instance_cells(m['ctr_hmp'], m['offsets'], up)
                    assert cells.dtype == torch.float32 and cells.dim() == 4
                    outs.append(e.postprocess(m['sem'], cells)[..., :size[0], :size[1]])
            for mo in list(e.median_queue)[e.mid_idx + 1:]:
                cells = e.get_instance_cells(mo['ctr_hmp'], mo['offsets'], up)
                outs.append(e.postprocess(mo['sem'], cells)[..., :size[0], :size[1]])
        assert len(outs) == sum(p['emitted']) + p['n_tail']
        for i, o in enumerate(outs):
            assert _mismatch(o, g[f'out_{i}']) == 0, f'{mode} slice {i}'


def _rle_equal(a, b):
    assert list(a.keys()) == list(b.keys())
    for c in a:
        assert list(a[c].keys()) == list(b[c].keys()), f'class {c}'
        for lab in a[c]:
            assert a[c][lab]['box'] == b[c][lab]['box']
            np.testing.assert_array_equal(a[c][lab]['starts'], b[c][lab]['starts'])
            np.testing.assert_array_equal(a[c][lab]['runs'], b[c][lab]['runs'])


@pytest.mark.parametrize('ks', [1, 3, 5])
def test_stack_shard_matches_sequential_engine(ks, cuda_device):
    """The z-sharded driver (deferred-sync block processing, inference/stack.py) must give, slice for
    slice, the RLE dicts of the sequential Render engine (median queue -> fused post-process -> RLE)."""
    from empanada_b200.inference import rle, stack
    from empanada_b200.synth import synth_stack_slices
    D, H, W = 11, 192, 256
    sl = list(synth_stack_slices(D, H, W, 40, seed=21, coarse=4, sigma=4.0, z_extent=(4, 12), semi_axes=(6, 20)))
    heads = [{k: torch.from_numpy(s[k]).to(cuda_device) for k in ('sem_prob', 'ctr_hmp', 'offsets')} for s in sl]

    class Probs(torch.nn.Module):                    # a "model" whose sem_logits are logit(prob)
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1, device=cuda_device))
            self.i = 0

        def forward(self, image, render_steps=None, interpolate_ins=None):
            h = heads[self.i]
            self.i += 1
            return {'sem_logits': torch.logit(h['sem_prob'].double()).float(), 'ctr_hmp': h['ctr_hmp'].clone(),
                    'offsets': h['offsets'].clone()}

    kw = dict(thing_list=[1], label_divisor=20000, stuff_area=64, void_label=0, nms_threshold=0.1, nms_kernel=3,
              confidence_thr=0.3, coarse_boundaries=True)
    seq = eng.PanopticDeepLabRenderEngine3d(Probs(), median_kernel_size=ks, **kw)
    pans = []
    for z in range(D):
        o = seq(torch.zeros(1, 1, H, W), (H, W), 1)
        if o is not None:
            pans.append(o)
    pans += seq.end(1)
    assert len(pans) == D
    want = [rle.pan_seg_to_rle_seg(p, [1], 20000, [1], True) for p in pans]

    # the shard must see exactly the probabilities the engine saw: sigmoid(logit(p)) in fp32
    probs = [eng.logits_to_prob(torch.logit(h['sem_prob'].double()).float()) for h in heads]
    e = eng.PanopticDeepLabRenderEngine(torch.nn.Identity(), **kw)
    shard = stack.StackShard(e, labels=[1], depth=D, rank=0, world_size=1, median_kernel_size=ks)
    for z in shard.slices():
        shard.add(z, probs[z], heads[z]['ctr_hmp'], heads[z]['offsets'], size=(H, W))
    got = shard.finish()
    assert sorted(got.keys()) == list(range(D))
    assert sum(len(v[1]) for v in got.values()) > 20
    for z in range(D):
        _rle_equal(got[z], want[z])

    # cross-slice matching of the block: one overlap launch + host chain == the dict-API chain slice by slice
    if ks == 3:
        from empanada_b200.inference import matcher as mt
        matched = shard.match(got)
        mat = mt.RLEMatcher(1, 20000, 0.25, 0.25, True)
        fwd = []
        for z in range(D):
            seg = got[z][1]
            if mat.target_rle is None:
                mat.initialize_target(seg)
            else:
                seg = mat(seg)
            fwd.append(seg)
        mat.target_rle, mat.assign_new = None, False
        for z in range(D - 1, -1, -1):
            seg = fwd[z]
            if mat.target_rle is None:
                mat.initialize_target(seg)
            else:
                seg = mat(seg)
            _rle_equal({1: matched[z][1]}, {1: seg})
        vol = shard.fill(torch.int64).cpu().numpy()
        for z in range(D):
            np.testing.assert_array_equal(vol[z], rle.rle_seg_to_pan_seg(matched[z], (H, W)).astype(np.int64))
        labels0 = set(matched[0][1].keys())
        assert any(set(matched[z][1].keys()) & labels0 for z in range(1, D))      # objects are tracked across slices
