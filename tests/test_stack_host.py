"""CPU: host-side logic of the z-sharded stack path — slice partitioning, the recursive median
chain with its rank-to-rank carry, and the label-offset all-gather — with world_size-2/3 gloo
process groups.  The sharded result must equal the sequential _MedianQueue semantics (oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from empanada_b200.inference import stack


def test_partition_covers_everything():
    for depth in (1, 7, 8, 512, 513):
        for world in (1, 2, 3, 8):
            blocks = [stack.partition_slices(depth, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == depth
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert stack.halo_range(512, 8, 0, 3) == (0, 65)
    assert stack.halo_range(512, 8, 7, 3) == (448, 512)
    assert stack.halo_range(512, 8, 3, 1) == (192, 256)


def _sequential(planes, ks):
    """What the reference's queue emits for every slice (engines.py:68-90), via the oracle."""
    q = oracle.MedianQueue(ks)
    out = []
    for p in planes:
        q.enqueue({'sem': p})
        o = q.get_next(['sem'])
        if o is not None:
            out.append(o['sem'])
    out += [e['sem'] for e in q.end()]
    return out


def _tmedian(window):
    return torch.median(torch.stack(window), dim=0).values


@pytest.mark.parametrize('ks', [1, 3, 5, 7])
@pytest.mark.parametrize('world', [1, 2, 3])
def test_median_chain_single_process(ks, world):
    rng = np.random.default_rng(ks * 10 + world)
    D = 13
    planes = [rng.random((1, 2, 5, 6), dtype=np.float32) for _ in range(D)]
    want = _sequential([p.copy() for p in planes], ks)
    assert len(want) == D
    carry = []
    for r in range(world):
        z0, z1 = stack.partition_slices(D, world, r)
        _, zh = stack.halo_range(D, world, r, ks)
        raw = {z: torch.from_numpy(planes[z]) for z in range(z0, zh)}
        got, carry = stack.median_chain(raw, z0, z1, D, ks, carry, _tmedian)
        for z in range(z0, z1):
            np.testing.assert_array_equal(got[z].numpy(), want[z])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ks, D, seed, sticky, ret):
    """One rank of the guessed-carry hand-over (stack.carry_rounds) with torch restatements of the two kernels: chain
    from a carry, and "repair" as a full re-run that reports whether the outgoing carry moved."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        planes = _planes(D, seed, sticky)                                            # same on all ranks
        z0, z1 = stack.partition_slices(D, world, rank)
        _, zh = stack.halo_range(D, world, rank, ks)
        raw = {z: torch.from_numpy(planes[z]) for z in range(z0, zh)}
        mid = (ks - 1) // 2
        st = {'got': None, 'out': [torch.zeros_like(raw[z0]) for _ in range(mid)], 'recv': [torch.zeros_like(raw[z0]) for _ in range(mid)],
              'rounds': 0}

        def run(carry):
            got, nxt = stack.median_chain(raw, z0, z1, D, ks, carry, _tmedian)
            moved = any(not torch.equal(a, b) for a, b in zip(st['out'], nxt))
            for a, b in zip(st['out'], nxt):
                a.copy_(b)
            st['got'] = got
            return moved

        def repair(rnd):
            st['rounds'] += 1
            return torch.tensor([int(run([t.clone() for t in st['recv']]))])

        def any_changed(flag):
            t = flag.clone() if flag is not None else torch.zeros(1, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return bool(t.item())

        for settle in (False, True):
            flag = stack.carry_rounds(rank, world, lambda: run([raw[z0]] * mid), repair,
                                      lambda: stack.exchange_planes(st['out'], st['recv'], rank, world),
                                      any_changed if settle else None)
            if not settle:
                first_flag = int(flag.item()) if flag is not None else 0
                one_round = {z: st['got'][z].numpy().copy() for z in st['got']}
        counts = torch.tensor([10 * (rank + 1), 3 + rank], dtype=torch.int64)
        offs, table = stack.label_offsets(counts)
        ret[rank] = ({z: st['got'][z].numpy() for z in st['got']}, offs.tolist(), table.tolist(), first_flag, one_round)
    finally:
        dist.destroy_process_group()


def _planes(D, seed, sticky):
    rng = np.random.default_rng(seed)
    if sticky:      # alternating 0 / 10 planes: every window's middle values are carried ones, the filter never forgets its start
        return [np.full((1, 1, 4, 8), 10.0 * (z % 2), np.float32) + (rng.random((1, 1, 4, 8), dtype=np.float32) if z == 1 else 0)
                for z in range(D)]
    return [rng.random((1, 1, 4, 8), dtype=np.float32) for _ in range(D)]


@pytest.mark.parametrize('world,ks,sticky', [(2, 3, False), (2, 5, False), (3, 3, False), (3, 3, True), (4, 5, True)])
def test_guessed_carry_rounds_and_offsets_gloo(world, ks, sticky):
    """Every rank starts its chain from a guessed carry, one exchange + repair makes all blocks exact unless some
    rank reports that its outgoing carry moved; the settle rounds then converge to the sequential queue's result."""
    D, seed = 13, 5
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ks, D, seed, sticky, ret), nprocs=world, join=True)
    want = _sequential(_planes(D, seed, sticky), ks)
    seen = set()
    flags = [ret[r][3] for r in range(world)]
    for r in range(world):
        got, offs, table, _, one_round = ret[r]
        for z, p in got.items():
            np.testing.assert_array_equal(p, want[z])
            if not any(flags):                       # nobody's carry moved: the single round was already exact
                np.testing.assert_array_equal(one_round[z], want[z])
            seen.add(z)
        assert offs == [sum(10 * (q + 1) for q in range(r)), sum(3 + q for q in range(r))]
        assert table == [[10 * (q + 1), 3 + q] for q in range(world)]
    assert seen == set(range(D))
    if sticky and world > 2:
        assert any(flags[1:])                        # the case the settle rounds exist for


def test_shard_rejects_blocks_shorter_than_the_window():
    """depth < world_size * mid cannot be sharded; every rank sees the same arguments, so every rank raises before any
    collective (nobody is left waiting)."""
    with pytest.raises(ValueError):
        stack.StackShard(None, [1], depth=6, rank=1, world_size=4, median_kernel_size=7)
    with pytest.raises(ValueError):
        stack.StackShard(None, [1], depth=2, rank=0, world_size=1, median_kernel_size=3)


def test_apply_label_offset():
    seg = {1: {1001: {'box': (0, 0, 1, 1)}, 1002: {'box': (1, 1, 2, 2)}}, 2: {2000: {'box': (0, 0, 4, 4)}}}
    out = stack.apply_label_offset(seg, {1: 40, 2: 7}, 1000, [1])
    assert list(out[1]) == [1041, 1042] and list(out[2]) == [2000]
    with pytest.raises(ValueError):
        stack.apply_label_offset(seg, {1: 998}, 1000, [1])


# ---- cross-slice matching across block boundaries (host chain; the overlaps come from the oracle here) ----
@pytest.mark.parametrize('name', ['matcher_stack_a', 'matcher_stack_b', 'matcher_stack_nofc'])
def test_matching_chain_continues_across_blocks(name):
    """StackMatcher.forward/backward with the state handed over between two z-blocks (what two ranks
    exchange) must equal the single-block chain, i.e. the reference's RLEMatcher loops
    (tests/golden/matcher_stack_*.npz were produced by the unmodified reference)."""
    import oracle
    from conftest import load_golden
    from oracle import matcher as om
    from empanada_b200.inference import matcher as mt
    g = load_golden(name)
    p = g['params']
    D = p['D']
    rles = [oracle.pan_seg_to_rle_seg(g['in_vol'][z], [1], 1000, [1], p['force_connected'])[1] for z in range(D)]
    ov = []
    for a, b in zip(rles[:-1], rles[1:]):
        m = np.array([[om.rle_intersection(x['starts'], x['runs'], y['starts'], y['runs']) for y in b.values()] for x in a.values()],
                     np.int64).reshape(len(a), len(b))
        i, j = np.nonzero(m)
        ov.append((i, j, m[i, j]))
    for cut in (1, D // 2, D - 1):
        A, B = rles[:cut], rles[cut:]
        sm_a, sm_b = mt.StackMatcher(1, 1000, 0.25, 0.25), mt.StackMatcher(1, 1000, 0.25, 0.25)
        f_a, g_a = sm_a.forward(A, ov[:cut - 1])
        f_b, g_b = sm_b.forward(B, ov[cut:], prev=dict(sm_a.forward_state(g_a), overlaps=ov[cut - 1]))
        assert sm_b.matcher.next_label == int(g['fwd_next_label'])
        b_b = sm_b.backward(f_b, g_b, B, ov[cut:])
        b_a = sm_a.backward(f_a, g_a, A, ov[:cut - 1], nxt=dict(sm_b.backward_state(b_b), overlaps=ov[cut - 1]))
        for z, got in enumerate(f_a + f_b):
            np.testing.assert_array_equal(oracle.rle_seg_to_pan_seg({1: got}, (p['H'], p['W'])), g[f'fwd_{z}'])
        for z, got in enumerate(b_a + b_b):
            np.testing.assert_array_equal(oracle.rle_seg_to_pan_seg({1: got}, (p['H'], p['W'])), g[f'bwd_{z}'])


# ---- host side of the batched stack path: packed tables -> RleStack (no GPU needed) ------------------------------
def _packed_block(slices):
    """Packs per-slice instance lists [(class, label, box, starts, lens), ...] the way emp_stack_block does
    (include/empanada_b200.h, EMP_BLK_*): header | per-slice rows | int32 starts | int32 lengths | int64 instance rows."""
    from empanada_b200 import _cabi as C
    B = len(slices)
    starts, lens, inst, rows = [], [], [], []
    for sl in slices:
        first_inst, first_run = len(inst), len(starts)
        for cls, lab, box, s, l in sl:
            inst.append([cls, lab, *box, len(s), len(starts), int(np.sum(l))])
            starts += list(s)
            lens += list(l)
        rows.append([len(sl), first_inst, first_run, len(starts) - first_run, 0, len(sl)])
    R, Rp = len(starts), (len(starts) + 1) & ~1
    words = np.zeros(C.BLK_HDR_WORDS + C.BLK_SLICE_WORDS * B + Rp + C.BLK_INST_WORDS * len(inst), np.int64)
    words[0], words[1], words[2], words[3] = B, R, len(inst), C.BLK_INST_WORDS
    s0 = C.BLK_HDR_WORDS + C.BLK_SLICE_WORDS * B
    words[C.BLK_HDR_WORDS:s0] = np.asarray(rows, np.int64).reshape(-1)
    r32 = words[s0:s0 + Rp].view(np.int32)
    r32[:R] = starts
    r32[Rp:Rp + R] = lens
    words[s0 + Rp:] = np.asarray(inst, np.int64).reshape(-1)
    return words


def test_rle_stack_reads_packed_tables():
    slices = [
        [(1, 20001, (0, 0, 2, 9), [0, 12], [3, 2]), (1, 20002, (3, 1, 4, 4), [31], [3]), (2, 40000, (0, 5, 6, 10), [5, 15, 25], [5, 5, 5])],
        [],
        [(1, 20001, (1, 1, 2, 2), [11], [1])],
    ]
    words = _packed_block(slices)
    assert stack._BlockTables.words_needed(words, 3) == words.size
    t = stack._BlockTables(words, 3)
    out = stack.RleStack([1, 2], [1], 20000)
    for b, z in enumerate((7, 8, 9)):
        out._add(z, t, b)
    out._add_dict(10, {1: {20003: {'box': (0, 0, 1, 1), 'starts': np.array([0]), 'runs': np.array([1])}}, 2: {}})
    out.offsets = {1: 40, 2: 0}                         # this rank's label offset: thing classes only
    assert list(out) == [7, 8, 9, 10] and len(out) == 4
    assert out.counts() == (5, 8)
    seg = out[7]
    assert list(seg) == [1, 2] and list(seg[1]) == [20041, 20042] and list(seg[2]) == [40000]
    assert seg[1][20041]['box'] == (0, 0, 2, 9)
    np.testing.assert_array_equal(seg[1][20041]['starts'], [0, 12])
    np.testing.assert_array_equal(seg[1][20041]['runs'], [3, 2])
    np.testing.assert_array_equal(seg[2][40000]['starts'], [5, 15, 25])
    assert seg[1][20041]['starts'].dtype == np.int64 and seg[1][20041]['runs'].dtype == np.int64
    assert out[8] == {1: {}, 2: {}}
    assert list(out[9][1]) == [20041] and list(out[10][1]) == [20043]
    assert out[7] is seg                                # materialised once
    assert out.inst_rows(10) is None and out.inst_rows(7).shape == (3, 9)
    out.offsets = {1: 19999}
    out._cache.clear()
    with pytest.raises(ValueError):
        out[7]                                          # 20001 + 19999 leaves class 1's label range


def test_pinned_pool_hands_out_idle_buffers_only():
    """The pool looks at reference counts: a buffer whose views are still alive is not handed out again."""
    if not torch.cuda.is_available():
        pytest.skip('pinned memory needs a CUDA runtime')
    pool = stack._PinnedPool()
    dev = torch.device('cuda', 0)
    a = pool.acquire(dev, 1000)
    view = a.numpy()[:10]
    del a
    b = pool.acquire(dev, 500)
    assert b.data_ptr() != view.ctypes.data             # the first buffer is still referenced by `view`
    del view, b
    c = pool.acquire(dev, 800)
    assert len(pool.buffers[0]) == 2 and c.numel() >= 800


def test_shard_has_no_cpu_path_and_streams_nothing_before_a_slice_is_in():
    """finish() and advance() on CPU tensors fail loudly (no fallback); advance() before any add() is a no-op."""
    import torch
    from empanada_b200.inference import engines as eng
    e = eng.PanopticDeepLabRenderEngine(torch.nn.Identity(), thing_list=[1], label_divisor=20000, nms_kernel=3, confidence_thr=0.3)
    D, H, W = 6, 32, 48
    shard = stack.StackShard(e, labels=[1], depth=D, median_kernel_size=3, block=2)
    assert shard.advance() == 0
    for z in shard.slices():
        shard.add(z, torch.rand(1, 1, H, W), torch.rand(1, 1, H // 4, W // 4), torch.zeros(1, 2, H // 4, W // 4), size=(H, W))
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        shard.advance()
    shard = stack.StackShard(e, labels=[1], depth=D, median_kernel_size=3)
    for z in shard.slices():
        shard.add(z, torch.rand(1, 1, H, W), torch.rand(1, 1, H // 4, W // 4), torch.zeros(1, 2, H // 4, W // 4), size=(H, W))
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        shard.finish()
