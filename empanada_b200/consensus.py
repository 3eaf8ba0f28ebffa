"""
Drop-in for the tracker part of ``empanada.consensus`` (reference empanada/consensus.py):
``merge_objects_from_trackers`` and ``merge_semantic_from_trackers`` — the orthoplane consensus that
fuses the 3D instances found along xy / xz / yz (or any number of trackers) into one segmentation.

Where the work goes, and where it runs here:
  * the object graph (consensus.py:233-287) needs the voxel intersection of every pair of objects from
    different trackers whose boxes overlap — per pair a concatenate + argsort + sweep of two run lists in
    the reference.  Here: the runs of each tracker are flattened once and ONE launch per tracker pair
    (libempanada_b200 ``emp_rle_list_overlaps``) yields every non-zero intersection;
  * voting (array_utils.vote_by_ranges / rle_voting, array_utils.py:539-617): the indices covered by at
    least ``vote_thr`` ranges — a +1/-1 event sweep in array arithmetic instead of per-index vote lists;
  * the cluster logic (consensus.py:10-142: connected components, average-linkage cluster graph,
    iterative push / pull merging) is small-graph bookkeeping and stays on networkx, built and traversed
    in the reference's order so that labels and merge decisions come out identical.

The tile-merging functions of the reference (``merge_*_from_tiles``) belong to 2D tiled inference and
are not part of this path.
"""
import ctypes
from itertools import combinations

import networkx as nx
import numpy as np
import torch

from empanada_b200 import _cabi as C
from empanada_b200.inference.matcher import merge_boxes, merge_rles

__all__ = ['merge_objects_from_trackers', 'merge_semantic_from_trackers', 'merge_instances', 'vote_by_ranges',
           'join_ranges', 'object_overlaps']

MIN_OVERLAP = 100       # consensus.py:7-8
MIN_IOU = 1e-2


# ---- voting (array_utils.py:449-617, :665-671) ------------------------------------------------------
def _as_ranges(starts, runs):
    starts = np.asarray(starts, dtype=np.int64)
    return np.stack([starts, starts + np.asarray(runs, dtype=np.int64)], axis=1)


def _covered_at_least(list_of_ranges, k):
    """Maximal [start, end) intervals whose indices lie in at least k of the given ranges."""
    ranges = np.concatenate(list_of_ranges, axis=0)
    ranges = ranges[ranges[:, 1] > ranges[:, 0]]
    if ranges.shape[0] == 0:
        return np.zeros((0, 2), np.int64)
    pos = np.concatenate([ranges[:, 0], ranges[:, 1]])
    delta = np.concatenate([np.ones(ranges.shape[0], np.int64), -np.ones(ranges.shape[0], np.int64)])
    order = np.argsort(pos, kind='stable')
    pos, delta = pos[order], delta[order]
    uniq, first = np.unique(pos, return_index=True)
    depth = np.cumsum(np.add.reduceat(delta, first))                # coverage on [uniq[i], uniq[i+1])
    inside = depth >= k
    rise = inside & ~np.concatenate(([False], inside[:-1]))
    fall = ~inside & np.concatenate(([False], inside[:-1]))
    return np.stack([uniq[rise], uniq[fall]], axis=1)


def join_ranges(list_of_ranges):
    """Union of ranges; ranges that overlap or touch are joined (array_utils.py:634-671)."""
    list_of_ranges = [r for r in list_of_ranges if len(r) > 0]
    cat = np.concatenate(list_of_ranges, axis=0)
    s, r = merge_rles(cat[:, 0], cat[:, 1] - cat[:, 0])
    return np.stack([s, s + r], axis=1)


def vote_by_ranges(list_of_ranges, vote_thr=2):
    """Ranges of indices that at least vote_thr of the given range lists cover (array_utils.py:602-617).
    Returns an (m, 2) int64 array of [start, end) ranges — shape (0,) when nothing passes, like the
    reference's ``np.array([])``."""
    list_of_ranges = [r for r in list_of_ranges if len(r) > 0]
    if vote_thr == 1:
        return join_ranges(list_of_ranges)
    if len(list_of_ranges) < vote_thr:
        return np.array([])
    voted = _covered_at_least(list_of_ranges, vote_thr)
    return voted if voted.shape[0] else np.array([])


# ---- intersections ------------------------------------------------------------------------------------
def _intersection(starts_a, runs_a, starts_b, runs_b):
    """Voxels two run lists share (each list sorted and disjoint) — for the handful of instances inside
    one cluster (merge_overlapping); the bulk goes through object_overlaps on the GPU."""
    sa, sb = np.asarray(starts_a, np.int64), np.asarray(starts_b, np.int64)
    ea, eb = sa + np.asarray(runs_a, np.int64), sb + np.asarray(runs_b, np.int64)
    lo = np.searchsorted(ea, sb, side='right')                      # first a-run ending after each b-start
    hi = np.searchsorted(sa, eb, side='left')                       # a-runs starting before each b-end
    total = 0
    for j in np.flatnonzero(hi > lo):
        total += int((np.minimum(ea[lo[j]:hi[j]], eb[j]) - np.maximum(sa[lo[j]:hi[j]], sb[j])).sum())
    return total


def object_overlaps(tracker_indices, object_starts, object_runs, device=None):
    """Non-zero voxel intersections between objects of DIFFERENT trackers: arrays (i, j, inter) with
    i < j node indices, in lexicographic order — one emp_rle_list_overlaps launch per tracker pair."""
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    tracker_indices = np.asarray(tracker_indices)
    sides = {}
    for t in np.unique(tracker_indices):
        nodes = np.flatnonzero(tracker_indices == t)
        lens = [len(object_starts[n]) for n in nodes]
        if sum(lens) == 0:
            continue
        st = np.concatenate([np.asarray(object_starts[n], np.int64) for n in nodes])
        ru = np.concatenate([np.asarray(object_runs[n], np.int64) for n in nodes])
        slot = np.repeat(nodes.astype(np.int64), lens)
        order = np.argsort(st, kind='stable')
        table = np.ascontiguousarray(np.stack([st[order], ru[order], slot[order]], 1))
        sides[int(t)] = (torch.from_numpy(table).to(device), int(table.shape[0]), int(ru.max()))
    L = C.lib()
    rows = []
    for ta, tb in combinations(sorted(sides), 2):
        (A, na, lmax), (B, nb, _) = sides[ta], sides[tb]
        cap = max(4096, 4 * max(na, nb))
        while True:
            out = torch.empty((cap, 4), dtype=torch.int32, device=device)
            count = torch.zeros(1, dtype=torch.int32, device=device)
            with torch.cuda.device(device):
                C.check(L.emp_rle_list_overlaps(ctypes.c_void_p(A.data_ptr()), na, lmax, ctypes.c_void_p(B.data_ptr()), nb,
                                                ctypes.c_void_p(out.data_ptr()), cap, ctypes.c_void_p(count.data_ptr()),
                                                C.stream_ptr(device)))
            n = int(count.item())
            if n <= cap:
                break
            cap = n
        if n:
            rows.append(out[:n, 1:].to(torch.int64))
    if not rows:
        z = np.zeros(0, np.int64)
        return z, z, z
    rows = torch.cat(rows)
    lo, hi = torch.minimum(rows[:, 0], rows[:, 1]), torch.maximum(rows[:, 0], rows[:, 1])
    key = (lo << 31) | hi
    uniq, inv = torch.unique(key, return_inverse=True)              # sorted: lexicographic (i, j)
    inter = torch.zeros(uniq.shape[0], dtype=torch.int64, device=device).index_add_(0, inv, rows[:, 2])
    uniq, inter = uniq.cpu().numpy(), inter.cpu().numpy()
    return uniq >> 31, uniq & ((1 << 31) - 1), inter


# ---- cluster logic (consensus.py:10-195), on networkx like the reference --------------------------------
def _mean_link(G, cluster1, cluster2, key):
    """Average edge weight between two groups of nodes, absent edges counting 0 (consensus.py:10-33)."""
    total, n = 0, 0
    for u in cluster1:
        for v in cluster2:
            total = total + (G[u][v][key] if G.has_edge(u, v) else 0)
            n += 1
    return total / n


def _cluster_graph(G, cluster_iou_thr):
    """Nodes = connected groups of detections once edges with IoU <= thr are dropped; edges = groups
    whose average IoU / overlap is non-trivial (consensus.py:35-74)."""
    strong = G.copy()
    for u, v, d in G.edges(data=True):
        if d['iou'] <= cluster_iou_thr:
            strong.remove_edge(u, v)
    CG = nx.Graph()
    for i, members in enumerate(nx.connected_components(strong)):
        CG.add_node(i, cluster=members)
    for a, b in combinations(CG.nodes, 2):
        ca, cb = CG.nodes[a]['cluster'], CG.nodes[b]['cluster']
        iou_w = _mean_link(G, ca, cb, 'iou')
        ov_w = _mean_link(G, ca, cb, 'overlap')
        if iou_w > MIN_IOU or ov_w > MIN_OVERLAP:
            CG.add_edge(a, b, iou=iou_w, overlap=ov_w)
    return CG


def _absorb(H, src, dst):
    H.nodes[dst]['cluster'] = H.nodes[dst]['cluster'].union(H.nodes[src]['cluster'])
    H.remove_edge(src, dst)


def _merge_clusters(CG):
    """Iteratively resolve the cluster graph (consensus.py:86-142): take the most connected group; if a
    neighbour is bigger, dissolve it into its neighbours, else pull all neighbours (and their edges) in."""
    H = CG.copy()
    while H.number_of_edges() > 0:
        hub = sorted(H.nodes, key=lambda x: len(list(H.neighbors(x))), reverse=True)[0]
        around = sorted(H.neighbors(hub), key=lambda x: len(H.nodes[x]['cluster']), reverse=True)
        if len(H.nodes[around[0]]['cluster']) > len(H.nodes[hub]['cluster']):
            for nb in around:
                _absorb(H, hub, nb)
            H.remove_node(hub)
        else:
            for nb in around:
                _absorb(H, nb, hub)
                for far in list(H.neighbors(nb)):
                    if not H.has_edge(hub, far):
                        H.add_edge(hub, nb, iou=H[nb][far]['iou'])
                H.remove_node(nb)
    return H


def merge_instances(instances_dict):
    """Union of any number of instances: boxes united, runs joined (consensus.py:144-164)."""
    items = list(instances_dict.values())
    if len(items) < 2:
        return items[0]
    box, starts, runs = items[0]['box'], items[0]['starts'], items[0]['runs']
    for attrs in items[1:]:
        box = merge_boxes(box, attrs['box'])
        starts, runs = merge_rles(starts, runs, attrs['starts'], attrs['runs'])
    return dict(box=box, starts=starts, runs=runs)


def _merge_overlapping(cluster_instances):
    """Instances of one component that still share voxels non-trivially are united (consensus.py:166-195)."""
    if len(cluster_instances) < 2:
        return list(cluster_instances.values())
    ids = list(cluster_instances.keys())
    area = {k: int(np.sum(cluster_instances[k]['runs'])) for k in ids}
    MG = nx.Graph()
    MG.add_nodes_from(ids)
    for a, b in combinations(ids, 2):
        inter = _intersection(cluster_instances[a]['starts'], cluster_instances[a]['runs'],
                              cluster_instances[b]['starts'], cluster_instances[b]['runs'])
        if inter / (area[a] + area[b] - inter) > MIN_IOU or inter > MIN_OVERLAP:
            MG.add_edge(a, b)
    return [merge_instances({k: v for k, v in cluster_instances.items() if k in comp}) for comp in nx.connected_components(MG)]


# ---- the two entry points ---------------------------------------------------------------------------
def merge_semantic_from_trackers(semantic_trackers, pixel_vote_thr=2):
    """Voxel-wise vote over the (single) instance of each tracker (consensus.py:289-346).
    Returns {1: {'box', 'starts', 'runs'}} or {}."""
    boxes, ranges = [], []
    for tr in semantic_trackers:
        assert len(tr.instances.keys()) <= 1, 'Semantic classes only have 1 label!'
        for attrs in tr.instances.values():
            boxes.append(attrs['box'])
            ranges.append(_as_ranges(attrs['starts'], attrs['runs']))
    if not boxes:
        return {}
    box = boxes[0]
    for other in boxes[1:]:
        box = merge_boxes(box, other)
    voted = vote_by_ranges(ranges, pixel_vote_thr)
    return {1: {'box': box, 'starts': voted[:, 0], 'runs': voted[:, 1] - voted[:, 0]}}


def merge_objects_from_trackers(object_trackers, pixel_vote_thr=2, cluster_iou_thr=0.75, bypass=False):
    """The consensus algorithm over any number of trackers (consensus.py:348-469): objects from different
    trackers that overlap form a graph; per connected component the detections are grouped by IoU,
    groups resolved (``_merge_clusters``), each surviving group voted voxel by voxel, and instances of a
    component that still overlap are united.  Returns {instance id (1..n): {'box', 'starts', 'runs'}}."""
    n_votes = len(object_trackers)
    min_cluster_size = 1 if bypass else n_votes // 2 + 1
    if pixel_vote_thr < min_cluster_size:
        cluster_iou_thr = 0                                         # maximal merging without a majority vote

    source, boxes, starts, runs = [], [], [], []
    for t, tr in enumerate(object_trackers):
        for attrs in tr.instances.values():
            source.append(t)
            boxes.append(attrs['box'])
            starts.append(attrs['starts'])
            runs.append(attrs['runs'])
    if len(boxes) == 0:
        return {}

    # object graph: one node per detection, an edge wherever detections of different trackers intersect
    area = np.array([int(np.sum(r)) for r in runs], dtype=np.int64)
    G = nx.Graph()
    G.add_nodes_from(range(len(boxes)))
    ii, jj, inter = object_overlaps(np.array(source), starts, runs)
    for i, j, v in zip(ii.tolist(), jj.tolist(), inter.tolist()):
        G.add_edge(i, j, iou=v / (int(area[i]) + int(area[j]) - v), overlap=v)

    instances = {}
    next_id = 1
    for comp in nx.connected_components(G):
        if len(comp) < min_cluster_size:
            continue
        resolved = _merge_clusters(_cluster_graph(G.subgraph(comp), cluster_iou_thr))
        found = {}
        for node in resolved.nodes:
            members = list(resolved.nodes[node]['cluster'])
            if len(members) < min_cluster_size:
                continue
            box = boxes[members[0]]
            for m in members[1:]:
                box = merge_boxes(box, boxes[m])
            voted = vote_by_ranges([_as_ranges(starts[m], runs[m]) for m in members], pixel_vote_thr)
            if len(voted) > 0:
                found[len(found) + 1] = {'box': tuple(int(b) for b in box), 'starts': voted[:, 0], 'runs': voted[:, 1] - voted[:, 0]}
        for attrs in _merge_overlapping(found):
            instances[next_id] = attrs
            next_id += 1
    return instances
