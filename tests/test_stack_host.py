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


def _worker(rank, world, port, ks, D, seed, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        planes = [rng.random((1, 1, 4, 8), dtype=np.float32) for _ in range(D)]      # same on all ranks
        z0, z1 = stack.partition_slices(D, world, rank)
        _, zh = stack.halo_range(D, world, rank, ks)
        raw = {z: torch.from_numpy(planes[z]) for z in range(z0, zh)}
        mid = (ks - 1) // 2
        got = stack.exchange_carry(lambda c: stack.median_chain(raw, z0, z1, D, ks, c, _tmedian),
                                   rank, world, mid, raw[z0])
        counts = torch.tensor([10 * (rank + 1), 3 + rank], dtype=torch.int64)
        offs, table = stack.label_offsets(counts)
        ret[rank] = ({z: got[z].numpy() for z in got}, offs.tolist(), table.tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,ks', [(2, 3), (2, 5), (3, 3)])
def test_sharded_chain_and_offsets_gloo(world, ks):
    D, seed = 11, 5
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ks, D, seed, ret), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    planes = [rng.random((1, 1, 4, 8), dtype=np.float32) for _ in range(D)]
    want = _sequential(planes, ks)
    seen = set()
    for r in range(world):
        got, offs, table = ret[r]
        for z, p in got.items():
            np.testing.assert_array_equal(p, want[z])
            seen.add(z)
        assert offs == [sum(10 * (q + 1) for q in range(r)), sum(3 + q for q in range(r))]
        assert table == [[10 * (q + 1), 3 + q] for q in range(world)]
    assert seen == set(range(D))


# ---- ks == 3: the block as one clamp, the carry crossing the ranks without waiting for their chains ----
@pytest.mark.parametrize('world', [1, 2, 3, 5])
def test_compose_median3_equals_the_chain(world):
    """f_last = min(max(f_in, A), B) with (A, B) from the block's own raw planes must be the chain's last plane, for
    every block of every partition, fed with the true plane below it — including the raw first / last slices, equal
    neighbours (ties) and multi-channel planes."""
    rng = np.random.default_rng(100 + world)
    D = 17
    planes = [np.round(rng.random((1, 2, 5, 6), dtype=np.float32) * 8) / 8 for _ in range(D)]      # coarse values: many ties
    want = _sequential([p.copy() for p in planes], 3)
    for r in range(world):
        z0, z1 = stack.partition_slices(D, world, r)
        _, zh = stack.halo_range(D, world, r, 3)
        raw = {z: torch.from_numpy(planes[z]) for z in range(z0, zh)}
        A, B = stack.compose_median3(raw, z0, z1, D)
        below = torch.from_numpy(want[z0 - 1]) if z0 > 0 else torch.full_like(A, 123.0)     # rank 0: anything
        got = torch.minimum(torch.maximum(below, A), B)
        np.testing.assert_array_equal(got.numpy(), want[z1 - 1])


def _worker_median3(rank, world, port, D, seed, poison, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        planes = [rng.random((1, 1, 4, 8), dtype=np.float32) for _ in range(D)]      # same on all ranks
        if poison:
            planes[2][0, 0, 1, 3] = np.nan
        z0, z1 = stack.partition_slices(D, world, rank)
        _, zh = stack.halo_range(D, world, rank, 3)
        raw = {z: torch.from_numpy(planes[z]) for z in range(z0, zh)}
        got, mismatch = stack.exchange_carry_median3(lambda: stack.compose_median3(raw, z0, z1, D),
                                                     lambda c: stack.median_chain(raw, z0, z1, D, 3, c, _tmedian),
                                                     rank, world, raw[z0])
        ret[rank] = ({z: got[z].numpy() for z in got}, int(mismatch))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_median3_carry_crosses_ranks_as_clamps_gloo(world):
    D, seed = 11, 9
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_median3, args=(world, port, D, seed, False, ret), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    want = _sequential([rng.random((1, 1, 4, 8), dtype=np.float32) for _ in range(D)], 3)
    seen = set()
    for r in range(world):
        got, mismatch = ret[r]
        assert mismatch == 0
        for z, p in got.items():
            np.testing.assert_array_equal(p, want[z])
            seen.add(z)
    assert seen == set(range(D))


def test_median3_nan_is_flagged_gloo():
    """A NaN makes "composed clamp == chain" unprovable: the rank that handed such a plane on must say so (the
    driver then redoes the block with the sequential hand-over)."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_median3, args=(2, port, 11, 9, True, ret), nprocs=2, join=True)
    assert ret[0][1] == 1 and ret[1][1] == 0


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
