"""Orthoplane consensus against fixtures produced by the reference's own consensus.py
(tests/golden/consensus_*.npz: the sphere construction of the reference's tests/test_consensus.py and a
noisy 26-object case, five parameter settings each, plus the semantic vote).  The host logic is checked
on the CPU with the intersections computed by the oracle; the GPU test runs the real thing."""
import json

import numpy as np
import pytest
import torch

import oracle
from conftest import golden_names, load_golden
from empanada_b200 import consensus as cons
from empanada_b200.inference import tracker as tk


def _trackers(g, rle_fn):
    trs = []
    for i in range(3):
        vol = g[f'in_vol_{i}'].astype(np.int64)
        tr = tk.InstanceTracker(1, 1000, vol.shape, axis='xy')
        for z, sl in enumerate(vol):
            tr.update(rle_fn(np.ascontiguousarray(sl))[1], z)
        tr.finish()
        trs.append(tr)
    return trs


def _check(inst, g, prefix):
    labs = list(inst.keys())
    np.testing.assert_array_equal(np.asarray(labs, np.int64), g[f'{prefix}_labels'])
    if not labs:
        return
    np.testing.assert_array_equal(np.asarray([inst[l]['box'] for l in labs], np.int64), g[f'{prefix}_boxes'])
    np.testing.assert_array_equal(np.asarray([len(inst[l]['starts']) for l in labs]), g[f'{prefix}_counts'])
    np.testing.assert_array_equal(np.concatenate([inst[l]['starts'] for l in labs]), g[f'{prefix}_starts'])
    np.testing.assert_array_equal(np.concatenate([inst[l]['runs'] for l in labs]), g[f'{prefix}_runs'])


def _run_all(g, rle_fn):
    p = g['params']
    for k, (vote, thr, bypass) in enumerate(p['settings']):
        inst = cons.merge_objects_from_trackers(_trackers(g, rle_fn), pixel_vote_thr=vote, cluster_iou_thr=thr, bypass=bypass)
        _check(inst, g, f'obj{k}')
    for k, vote in enumerate(p['sem_votes']):
        trs = _trackers(g, rle_fn)
        for tr in trs:
            if tr.instances:
                tr.instances = {1001: cons.merge_instances(tr.instances)}
        if len(g[f'sem{k}_labels']) == 0:
            continue
        _check(cons.merge_semantic_from_trackers(trs, pixel_vote_thr=vote), g, f'sem{k}')


def _oracle_overlaps(tracker_indices, starts, runs, device=None):
    from oracle import matcher as om
    rows = []
    n = len(starts)
    for i in range(n):
        for j in range(i + 1, n):
            if tracker_indices[i] != tracker_indices[j]:
                v = om.rle_intersection(starts[i], runs[i], starts[j], runs[j])
                if v > 0:
                    rows.append((i, j, v))
    a = np.asarray(rows, np.int64).reshape(-1, 3)
    return a[:, 0], a[:, 1], a[:, 2]


@pytest.mark.parametrize('name', golden_names('consensus_'))
def test_consensus_host_logic(name, monkeypatch):
    monkeypatch.setattr(cons, 'object_overlaps', _oracle_overlaps)
    _run_all(load_golden(name), lambda sl: oracle.pan_seg_to_rle_seg(sl, [1], 1000, [1], False))


def test_vote_by_ranges_semantics():
    a = np.array([[0, 5], [10, 20], [30, 31]])
    b = np.array([[3, 12], [30, 31]])
    c = np.array([[4, 11], [19, 40]])
    np.testing.assert_array_equal(cons.vote_by_ranges([a, b, c], 2), [[3, 12], [19, 20], [30, 31]])
    np.testing.assert_array_equal(cons.vote_by_ranges([a, b, c], 3), [[4, 5], [10, 11], [30, 31]])
    np.testing.assert_array_equal(cons.vote_by_ranges([a, b, c], 1), [[0, 40]])
    np.testing.assert_array_equal(cons.vote_by_ranges([a, np.array([[5, 10]])], 1), [[0, 20], [30, 31]])      # touching ranges join
    assert cons.vote_by_ranges([a], 2).shape == (0,) and cons.vote_by_ranges([a, b], 3).shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize('name', golden_names('consensus_'))
def test_consensus_gpu(name, cuda_device):
    from empanada_b200.inference import rle
    g = load_golden(name)
    _run_all(g, lambda sl: rle.pan_seg_to_rle_seg(torch.from_numpy(sl).to(cuda_device), [1], 1000, [1], False))
    # the overlap launch on its own, against the oracle's pairwise intersections
    trs = _trackers(g, lambda sl: oracle.pan_seg_to_rle_seg(sl, [1], 1000, [1], False))
    src, st, ru = [], [], []
    for t, tr in enumerate(trs):
        for a in tr.instances.values():
            src.append(t)
            st.append(a['starts'])
            ru.append(a['runs'])
    got = cons.object_overlaps(np.array(src), st, ru)
    want = _oracle_overlaps(np.array(src), st, ru)
    for x, y in zip(got, want):
        np.testing.assert_array_equal(x, y)
