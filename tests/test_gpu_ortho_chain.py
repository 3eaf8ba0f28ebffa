"""B200: BASELINE configs[3] (orthoplane inference) in miniature, end to end through the reference-shaped
API — Render engine per slice along xy / xz / yz, RLE, forward + backward matching, trackers, filters,
instance consensus, filling — against tests/golden/ortho_chain.npz, which the reference's own
engines / rle / matcher / patterns / tracker / filters / consensus produced from the same head tensors
(tests/golden/make_golden.py: ortho_cases).  Everything is compared exactly: the per-axis label stacks,
each axis tracker's instances after filtering, the consensus instances and the consensus volume."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from empanada_b200.inference import engines as eng
from empanada_b200.inference import filters, patterns
from empanada_b200.inference import rle as erle

pytestmark = pytest.mark.gpu

AXES = {'xy': 0, 'xz': 1, 'yz': 2}


class ReplayModel(torch.nn.Module):
    def __init__(self, outputs, device):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1, device=device))
        self.outputs, self.i = outputs, 0

    def forward(self, image, render_steps=None, interpolate_ins=None):
        out = {k: v.clone() for k, v in self.outputs[self.i].items()}
        self.i += 1
        return out


def _check_instances(inst, g, prefix):
    labs = list(inst.keys())
    np.testing.assert_array_equal(np.asarray(labs, np.int64), g[f'{prefix}_labels'])
    np.testing.assert_array_equal(np.asarray([inst[l]['box'] for l in labs], np.int64).reshape(len(labs), -1),
                                  g[f'{prefix}_boxes'])
    np.testing.assert_array_equal(np.asarray([len(inst[l]['starts']) for l in labs], np.int64), g[f'{prefix}_counts'])
    np.testing.assert_array_equal(np.concatenate([np.asarray(inst[l]['starts']) for l in labs]), g[f'{prefix}_starts'])
    np.testing.assert_array_equal(np.concatenate([np.asarray(inst[l]['runs']) for l in labs]), g[f'{prefix}_runs'])


@pytest.fixture(scope='module')
def chain(cuda_device):
    """Runs the whole chain once; the tests below look at its stages."""
    g = load_golden('ortho_chain')
    P = g['params']
    shape, L = tuple(P['shape']), P['label_divisor']
    trackers = patterns.create_axis_trackers(AXES, P['labels'], L, shape)
    stacks = {}
    for name, axis in AXES.items():
        n = shape[axis]
        h, w = [d for a, d in enumerate(shape) if a != axis]
        outs = [{k: torch.from_numpy(g[f'in_{name}_{i}_{k}']).to(cuda_device) for k in ('sem_logits', 'ctr_hmp', 'offsets')}
                for i in range(n)]
        engine = eng.PanopticDeepLabRenderEngine3d(
            ReplayModel(outs, cuda_device), thing_list=P['thing_list'], median_kernel_size=P['median_kernel_size'],
            label_divisor=L, stuff_area=P['stuff_area'], void_label=P['void_label'], nms_threshold=P['nms_threshold'],
            nms_kernel=P['nms_kernel'], confidence_thr=P['confidence_thr'], padding_factor=P['padding_factor'],
            coarse_boundaries=P['coarse_boundaries'])
        matchers = patterns.create_matchers(P['thing_list'], L, P['merge_iou_thr'], P['merge_ioa_thr'])
        rle_stack = []

        def take(pan):
            # the CUDA tensor goes straight into the encoder (the script's .cpu().numpy() hop is optional here)
            seg = erle.pan_seg_to_rle_seg(pan.squeeze(), P['labels'], L, P['thing_list'], force_connected=True)
            rle_stack.append(patterns.apply_matchers(seg, matchers))

        for i in range(n):
            pan = engine(torch.zeros(1, 1, h, w), (h, w), upsampling=1)
            if pan is not None:
                take(pan)
        for pan in engine.end(1):
            take(pan)
        assert len(rle_stack) == n
        for index, seg in patterns.backward_matching(rle_stack, matchers, n):
            patterns.update_trackers(seg, index, trackers[name])
        patterns.finish_tracking(trackers[name])
        stack = np.zeros(shape, np.uint32)
        patterns.fill_panoptic_volume(stack, trackers[name])
        stacks[name] = stack
        for tr in trackers[name]:
            patterns.apply_filters(tr, [{'name': 'remove_small_objects', 'min_size': P['min_size']},
                                        {'name': 'remove_pancakes', 'min_span': P['min_span']}])
    cons = patterns.create_instance_consensus(patterns.get_axis_trackers_by_class(trackers, 1), P['pixel_vote_thr'],
                                              P['cluster_iou_thr'], False)
    filters.remove_small_objects(cons, min_size=P['min_size'])
    filters.remove_pancakes(cons, min_span=P['min_span'])
    return g, trackers, stacks, cons


@pytest.mark.parametrize('name', list(AXES))
def test_axis_stack_and_tracker(chain, name):
    g, trackers, stacks, _ = chain
    want = g[f'out_{name}_stack']
    assert stacks[name].dtype == want.dtype
    assert int((stacks[name] != want).sum()) == 0
    _check_instances(trackers[name][0].instances, g, f'out_{name}')


def test_consensus_instances_and_volume(chain, cuda_device):
    g, _, _, cons = chain
    _check_instances(cons.instances, g, 'out_cons')
    want = g['out_cons_vol']
    host = np.zeros(want.shape, np.uint32)
    patterns.fill_volume(host, cons.instances)                       # numpy volume: through HBM and back
    assert int((host != want).sum()) == 0
    dev = torch.zeros(want.shape, dtype=torch.int64, device=cuda_device)
    patterns.fill_volume(dev, cons.instances)                        # CUDA volume: painted in place
    assert int((dev.cpu().numpy() != want.astype(np.int64)).sum()) == 0
    small = np.zeros(want.shape, np.uint8)                           # the script's dtype for stuff classes
    attrs = next(iter(cons.instances.values()))
    patterns.fill_volume(small, {7: attrs})
    mask = np.zeros(small.size, bool)
    for s0, r in zip(attrs['starts'], attrs['runs']):
        mask[s0:s0 + r] = True
    np.testing.assert_array_equal(small.ravel() == 7, mask)
    assert set(np.unique(small)) <= {0, 7}


def test_multigpu_worker_loop(cuda_device):
    """patterns.forward_multigpu (the multi-GPU script's worker: median window, hardening, merge, RLE, forward
    matching on (sem, cells) pairs from a queue) gives the same matched stack as the engine path on the same
    (uncropped) head tensors."""
    import queue as pyqueue
    g = load_golden('ortho_chain')
    P = g['params']
    L, n = P['label_divisor'], P['shape'][0]
    heads = [{k: torch.from_numpy(g[f'in_xy_{i}_{k}']).to(cuda_device) for k in ('sem_logits', 'ctr_hmp', 'offsets')}
             for i in range(n)]
    Hp, Wp = heads[0]['ctr_hmp'].shape[-2:]
    kw = dict(thing_list=P['thing_list'], label_divisor=L, stuff_area=P['stuff_area'], void_label=P['void_label'],
              nms_threshold=P['nms_threshold'], nms_kernel=P['nms_kernel'], confidence_thr=P['confidence_thr'],
              coarse_boundaries=False)

    def matched_stack_via_engine():
        engine = eng.PanopticDeepLabRenderEngine3d(ReplayModel(heads, cuda_device), median_kernel_size=P['median_kernel_size'],
                                                   padding_factor=P['padding_factor'], **kw)
        matchers = patterns.create_matchers(P['thing_list'], L, P['merge_iou_thr'], P['merge_ioa_thr'])
        pans = [engine(torch.zeros(1, 1, Hp, Wp), (Hp, Wp), upsampling=1) for _ in range(n)]
        pans = [p for p in pans if p is not None] + engine.end(1)
        return [patterns.apply_matchers(erle.pan_seg_to_rle_seg(p.squeeze(), P['labels'], L, P['thing_list']), matchers)
                for p in pans]

    def matched_stack_via_worker():
        e2d = eng.PanopticDeepLabRenderEngine(torch.nn.Identity(), **kw)
        q = pyqueue.Queue()
        for hd in heads:
            q.put((torch.sigmoid(hd['sem_logits']), e2d.get_instance_cells(hd['ctr_hmp'], hd['offsets'], 1)))
        q.put(('finish', None))

        class Pipe:
            def send(self, obj):
                self.got = obj

            def close(self):
                pass

        pipe = Pipe()
        matchers = patterns.create_matchers(P['thing_list'], L, P['merge_iou_thr'], P['merge_ioa_thr'])
        patterns.forward_multigpu(matchers, q, [], pipe, P['confidence_thr'], P['median_kernel_size'], P['labels'], L,
                                  P['thing_list'], P['stuff_area'], P['void_label'])
        return pipe.got[0]

    a, b = matched_stack_via_engine(), matched_stack_via_worker()
    assert len(a) == len(b) == n
    for sa, sb in zip(a, b):
        assert list(sa.keys()) == list(sb.keys())
        for c in sa:
            assert list(sa[c].keys()) == list(sb[c].keys())
            for lab in sa[c]:
                assert tuple(sa[c][lab]['box']) == tuple(sb[c][lab]['box'])
                np.testing.assert_array_equal(sa[c][lab]['starts'], sb[c][lab]['starts'])
                np.testing.assert_array_equal(sa[c][lab]['runs'], sb[c][lab]['runs'])



class _FakeZarr:
    """The slice of zarr.Array's interface fill_volume uses: shape, dtype, chunks, z-slab reads and writes."""

    def __init__(self, shape, dtype, chunks):
        self.a = np.zeros(shape, dtype)
        self.shape, self.dtype, self.chunks = shape, np.dtype(dtype), chunks
        self.writes = []

    def __getitem__(self, k):
        return self.a[k]

    def __setitem__(self, k, v):
        self.writes.append(k)
        self.a[k] = v


@pytest.mark.parametrize('dtype', [np.uint32, np.uint16, np.int64])
def test_fill_volume_slabs_numpy_and_zarr_like(dtype, cuda_device, monkeypatch):
    """fill_volume on host volumes goes slab by slab through HBM (array_utils.numpy_fill_instances :725-736 /
    zarr_utils.zarr_fill_instances :88-175): same result as the reference's flat fill, runs that cross slab borders
    are split, runs past the end of the volume are clipped like numpy slicing clips them."""
    from empanada_b200.inference import fill as fl, patterns
    rng = np.random.default_rng(4)
    shape = (11, 24, 40)
    n = int(np.prod(shape))
    instances = {}
    for lab in range(1, 30):
        starts = np.sort(rng.choice(n - 400, size=12, replace=False)).astype(np.int64)
        runs = rng.integers(1, 300, size=12).astype(np.int64)            # many cross plane (and slab) borders
        instances[lab] = {'starts': starts, 'runs': runs}
    instances[99] = {'starts': np.array([n - 7], np.int64), 'runs': np.array([50], np.int64)}     # runs off the end
    want = np.zeros(n, dtype)
    for lab, a in instances.items():
        for s0, r in zip(a['starts'], a['runs']):
            want[s0:s0 + r] = lab
    want = want.reshape(shape)
    monkeypatch.setattr(fl.fill_slabs, '__defaults__', (3 * 24 * 40 * 4, None))       # 3 planes per slab
    vol = np.zeros(shape, dtype)
    patterns.fill_volume(vol, instances)
    np.testing.assert_array_equal(vol, want)
    z = _FakeZarr(shape, dtype, (2, 8, 8))
    patterns.fill_volume(z, instances, processes=3)
    np.testing.assert_array_equal(z.a, want)
    assert all(k.start % 2 == 0 for k in z.writes)                          # whole z-chunks per slab


@pytest.mark.parametrize('dtype', [torch.uint8, torch.int32, torch.float32])
@pytest.mark.parametrize('shape', [(40, 50, 70), (33, 65, 31), (64, 64, 64)])
def test_take_slices_matches_indexing(dtype, shape, cuda_device):
    """emp_take_slices — array_utils.take (array_utils.py:6-23) for a batch of consecutive slices of an HBM-resident
    volume — against plain indexing, along all three axes, at batch boundaries and odd sizes."""
    from empanada_b200.inference.volume import take_slices
    g = torch.Generator(device='cpu')
    g.manual_seed(3)
    vol = (torch.rand(shape, generator=g) * 250).to(dtype).to(cuda_device)
    for axis in range(3):
        n_ax = shape[axis]
        for i0, n in ((0, n_ax), (3, 1), (n_ax - 7, 7), (5, min(33, n_ax - 5))):
            got = take_slices(vol, axis, i0, n)
            want = vol.narrow(axis, i0, n).movedim(axis, 0).contiguous()
            assert got.shape == want.shape and torch.equal(got, want), (axis, i0, n)
    with pytest.raises(IndexError):
        take_slices(vol, 1, shape[1] - 2, 3)
