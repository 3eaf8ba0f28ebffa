"""
Drop-in for the RLE part of ``empanada.inference.matcher`` (reference empanada/inference/matcher.py):
``rle_matcher`` and ``RLEMatcher`` with the reference's signatures, return values and label
bookkeeping.  The pixel intersections behind every IoU / IoA — the reference's per-pair sort-and-sweep
on the host (array_utils.rle_intersection) — come from one CUDA launch over the run tables
(libempanada_b200 ``emp_rle_pair_overlaps``); the Hungarian assignment
(scipy.optimize.linear_sum_assignment, as in the reference) and the n x m matrix logic stay on the host.

Two ways in:
  * the reference's dict API — ``rle_matcher(target_rles, match_rles, ...)`` / ``RLEMatcher`` — uploads
    the two run lists, one launch per call;
  * ``block_overlaps`` + ``StackMatcher`` for the z-sharded stack driver: the run tables of a whole
    z-block are already in HBM (inference/stack.py), so ONE launch yields the overlaps of every
    consecutive slice pair, and the forward / backward matching chains (patterns.py:68-112) then run on
    the host without touching run lists at all — merged instances are handled as groups of the
    original per-slice instances, whose intersections add up because instances of a slice are disjoint.

``fast_matcher`` (dense label maps + skimage regionprops) is not part of the RLE path and is not provided.
"""
import ctypes

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

from empanada_b200 import _cabi as C
from empanada_b200.inference.rle import unpack_rle_attrs

__all__ = ['rle_matcher', 'RLEMatcher', 'merge_attrs', 'merge_rles', 'merge_boxes', 'pair_overlaps', 'block_overlaps',
           'StackMatcher']


# ---- host helpers with the reference's semantics (array_utils.py:101-125, :634-718) ---------------
def merge_boxes(box1, box2):
    n = len(box1)
    ndim = n // 2
    return tuple(min(box1[i], box2[i]) if i < ndim else max(box1[i], box2[i]) for i in range(n))


def merge_rles(starts_a, runs_a, starts_b=None, runs_b=None):
    """Union of possibly overlapping run lists as sorted, non-overlapping runs; ranges that overlap or
    touch are joined (array_utils.py:690-718, _join_ranges :634-663)."""
    s = np.asarray(starts_a, dtype=np.int64)
    e = s + np.asarray(runs_a, dtype=np.int64)
    if starts_b is not None and runs_b is not None:
        sb = np.asarray(starts_b, dtype=np.int64)
        s = np.concatenate([s, sb])
        e = np.concatenate([e, sb + np.asarray(runs_b, dtype=np.int64)])
    order = np.argsort(s, kind='stable')
    s, e = s[order], e[order]
    reach = np.maximum.accumulate(e)
    head = np.ones(s.shape[0], dtype=bool)
    head[1:] = s[1:] > reach[:-1]                       # a new range starts where nothing before reaches it
    idx = np.flatnonzero(head)
    ends = np.maximum.reduceat(e, idx)
    return s[idx], ends - s[idx]


def merge_attrs(rle_attr1, rle_attr2):
    """matcher.py:14-29."""
    starts, runs = merge_rles(rle_attr1['starts'], rle_attr1['runs'], rle_attr2['starts'], rle_attr2['runs'])
    return {'box': merge_boxes(rle_attr1['box'], rle_attr2['box']), 'starts': starts, 'runs': runs}


# ---- overlaps on the GPU ----------------------------------------------------------------------------
def _overlap_rows(runs, n_runs_dev, max_runs, device):
    """emp_rle_pair_overlaps on a (n_slices, run_stride, 3) int64 CUDA tensor -> summed rows
    (pair, slot_a, slot_b, inter) as int64 numpy arrays."""
    n_slices, run_stride = int(runs.shape[0]), int(runs.shape[1])
    L = C.lib()
    cap = max(1024, 2 * max_runs * max(n_slices - 1, 1))
    while True:
        out = torch.empty((cap, 4), dtype=torch.int32, device=device)
        count = torch.zeros(1, dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            C.check(L.emp_rle_pair_overlaps(ctypes.c_void_p(runs.data_ptr()), run_stride, ctypes.c_void_p(n_runs_dev.data_ptr()),
                                            n_slices, int(max_runs), ctypes.c_void_p(out.data_ptr()), cap,
                                            ctypes.c_void_p(count.data_ptr()), C.stream_ptr(device)))
        n = int(count.item())
        if n <= cap:
            break
        cap = n
    rows = out[:n].to(torch.int64)
    if n == 0:
        z = np.zeros(0, np.int64)
        return z, z, z, z
    key = (rows[:, 0] << 42) | (rows[:, 1] << 21) | rows[:, 2]          # slots < 2^21
    uniq, inv = torch.unique(key, return_inverse=True)
    inter = torch.zeros(uniq.shape[0], dtype=torch.int64, device=device).index_add_(0, inv, rows[:, 3])
    uniq, inter = uniq.cpu().numpy(), inter.cpu().numpy()
    return uniq >> 42, (uniq >> 21) & ((1 << 21) - 1), uniq & ((1 << 21) - 1), inter


def _flat_runs(starts_list, runs_list):
    """All runs of one side as (start, length, slot) rows in ascending start order."""
    if len(starts_list) == 0:
        return np.zeros((0, 3), np.int64)
    slot = np.repeat(np.arange(len(starts_list), dtype=np.int64), [len(s) for s in starts_list])
    st = np.concatenate([np.asarray(s, np.int64) for s in starts_list]) if slot.size else np.zeros(0, np.int64)
    ru = np.concatenate([np.asarray(r, np.int64) for r in runs_list]) if slot.size else np.zeros(0, np.int64)
    order = np.argsort(st, kind='stable')
    return np.stack([st[order], ru[order], slot[order]], 1)


def _disjoint(table):
    """True if no two runs of the start-ordered (start, length, slot) table overlap."""
    if table.shape[0] < 2:
        return True
    ends = table[:, 0] + table[:, 1]
    return not bool((table[1:, 0] < np.maximum.accumulate(ends)[:-1]).any())


def _list_overlap_rows(a, b, device):
    """emp_rle_list_overlaps (instances of one list may overlap each other) -> summed (slot_a, slot_b, inter)."""
    L = C.lib()
    A, B = torch.from_numpy(np.ascontiguousarray(a)).to(device), torch.from_numpy(np.ascontiguousarray(b)).to(device)
    cap = max(4096, 4 * max(a.shape[0], b.shape[0]))
    while True:
        out = torch.empty((cap, 4), dtype=torch.int32, device=device)
        count = torch.zeros(1, dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            C.check(L.emp_rle_list_overlaps(ctypes.c_void_p(A.data_ptr()), int(a.shape[0]), int(a[:, 1].max()),
                                            ctypes.c_void_p(B.data_ptr()), int(b.shape[0]), ctypes.c_void_p(out.data_ptr()), cap,
                                            ctypes.c_void_p(count.data_ptr()), C.stream_ptr(device)))
        n = int(count.item())
        if n <= cap:
            break
        cap = n
    rows = out[:n].to(torch.int64)
    if n == 0:
        z = np.zeros(0, np.int64)
        return z, z, z
    key = (rows[:, 1] << 31) | rows[:, 2]
    uniq, inv = torch.unique(key, return_inverse=True)
    inter = torch.zeros(uniq.shape[0], dtype=torch.int64, device=device).index_add_(0, inv, rows[:, 3])
    uniq, inter = uniq.cpu().numpy(), inter.cpu().numpy()
    return uniq >> 31, uniq & ((1 << 31) - 1), inter


def pair_overlaps(target_starts, target_runs, match_starts, match_runs, device=None):
    """(n, m) int64 matrix of pixel intersections between two lists of run-length encoded instances.  The instances of
    one slice are mutually disjoint and take the pair kernel (binary search over the target's run ends); lists whose
    instances overlap each other — the dict API accepts anything — go through the list kernel, which does not assume
    disjointness (array_utils.rle_intersection treats every pair independently, :371-403)."""
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    n, m = len(target_starts), len(match_starts)
    inter = np.zeros((n, m), np.int64)
    if n == 0 or m == 0:
        return inter
    a, b = _flat_runs(target_starts, target_runs), _flat_runs(match_starts, match_runs)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return inter
    if not (_disjoint(a) and _disjoint(b)):
        sa, sb, ov = _list_overlap_rows(a, b, device)
        inter[sa, sb] = ov
        return inter
    stride = max(a.shape[0], b.shape[0], 1)
    both = np.zeros((2, stride, 3), np.int64)
    both[0, :a.shape[0]] = a
    both[1, :b.shape[0]] = b
    runs = torch.from_numpy(both).to(device)
    n_runs = torch.tensor([a.shape[0], b.shape[0]], dtype=torch.int32, device=device)
    _, sa, sb, ov = _overlap_rows(runs, n_runs, stride, device)
    inter[sa, sb] = ov
    return inter


def block_overlaps(runs_all, n_runs, device=None):
    """Overlaps of every consecutive slice pair of a z-block whose run tables are in HBM.
    runs_all (n_slices, run_stride, 3) int64 CUDA tensor (emp_rle's runs_out per slice, rows in
    ascending start order), n_runs host array of valid rows per slice.  Returns a list of n_slices-1
    tuples (slot_a, slot_b, inter) of int64 numpy arrays."""
    device = runs_all.device if device is None else device
    n_runs = np.asarray(n_runs, np.int64)
    n = int(runs_all.shape[0])
    nr = torch.from_numpy(n_runs.astype(np.int32)).to(device)
    pair, sa, sb, ov = _overlap_rows(runs_all, nr, int(n_runs.max()) if n else 0, device)
    cuts = np.searchsorted(pair, np.arange(1, max(n - 1, 1)))
    return [tuple(x) for x in zip(np.split(sa, cuts), np.split(sb, cuts), np.split(ov, cuts))][:max(n - 1, 0)]


# ---- the reference's matcher ------------------------------------------------------------------------
def _match_from_inter(target_labels, match_labels, inter, area_t, area_m, iou_thr, return_iou, return_ioa):
    """matcher.py:196-232 given the intersections: IoU in float64 ('float' in the reference), IoA in
    float32, Hungarian on the IoU matrix, threshold on the assigned pairs.  `inter` is a dense (n, m)
    int64 matrix or its non-zero entries as (rows, cols, values)."""
    n, m = len(target_labels), len(match_labels)
    if isinstance(inter, tuple):
        r, c, v = inter
    else:
        r, c = np.nonzero(inter)
        v = inter[r, c]
    v = np.asarray(v, dtype=np.int64)
    iou_matrix = np.zeros((n, m), dtype='float')
    ioa_matrix = np.zeros((n, m), dtype=np.float32)
    iou_matrix[r, c] = v / (area_t[r] + area_m[c] - v)
    ioa_matrix[r, c] = v / area_m[c]
    match_rows, match_cols = linear_sum_assignment(iou_matrix, maximize=True)
    if iou_thr is not None:
        iou_mask = iou_matrix[match_rows, match_cols] >= iou_thr
        match_rows, match_cols = match_rows[iou_mask], match_cols[iou_mask]
    matched_labels = (target_labels[match_rows], match_labels[match_cols])
    output = (matched_labels, [target_labels, match_labels], iou_matrix[(match_rows, match_cols)])
    if return_iou:
        output = output + (iou_matrix,)
    if return_ioa:
        output = output + (ioa_matrix,)
    return output


def rle_matcher(target_instance_rles, match_instance_rles, iou_thr=0.5, return_iou=False, return_ioa=False,
                inter=None, areas=None):
    r"""Performs Hungarian matching on run length encodings (matcher.py:136-232).

    Same arguments and return values as the reference.  Optional, for callers that already hold them
    (StackMatcher): ``inter`` — the pixel intersections in the dicts' key order, an (n, m) int64 matrix
    or (rows, cols, values) — skips the GPU launch; ``areas`` — (target areas, match areas) int64."""
    target_labels = np.array([int(k) for k in target_instance_rles.keys()])
    match_labels = np.array([int(k) for k in match_instance_rles.keys()])
    if len(target_labels) == 0 or len(match_labels) == 0:
        empty = np.array([])
        if return_ioa:
            return (empty, empty), (target_labels, match_labels), empty, empty
        return (empty, empty), (target_labels, match_labels), empty
    if inter is None:
        _, _, target_starts, target_runs = unpack_rle_attrs(target_instance_rles)
        _, _, match_starts, match_runs = unpack_rle_attrs(match_instance_rles)
        inter = pair_overlaps(target_starts, target_runs, match_starts, match_runs)
    if areas is None:
        areas = (instance_areas(target_instance_rles), instance_areas(match_instance_rles))
    return _match_from_inter(target_labels, match_labels, inter, areas[0], areas[1], iou_thr, return_iou, return_ioa)


def instance_areas(instance_rles):
    """Pixels per instance, in dict order."""
    return np.array([int(np.sum(a['runs'])) for a in instance_rles.values()], dtype=np.int64)


class RLEMatcher:
    r"""Carries instance labels from slice to slice of a stack (reference matcher.py:234-326; same
    constructor arguments, attributes ``next_label`` / ``target_rle`` / ``assign_new`` and call semantics).

    For every instance of the incoming slice, in dict order:
      1. Hungarian-matched to a target instance with IoU >= ``merge_iou_thr``  -> that target's label;
      2. else, if its best IoA over the targets is >= ``merge_ioa_thr`` (a false split) -> the label of
         the target holding that IoA (first maximum);
      3. else a fresh label (``assign_new``) or its own label.
    Instances that end up with the same label are merged (boxes united, runs joined)."""

    def __init__(self, class_id, label_divisor, merge_iou_thr=0.25, merge_ioa_thr=0.25, assign_new=True, **kwargs):
        self.class_id = class_id
        self.label_divisor = label_divisor
        self.merge_iou_thr = merge_iou_thr
        self.merge_ioa_thr = merge_ioa_thr
        self.assign_new = assign_new
        self.next_label = class_id * label_divisor + 1
        self.target_rle = None
        self.last_assignment = None         # [(incoming label, label it received)] of the last call, in dict order

    def initialize_target(self, target_instance_rles):
        self.target_rle = target_instance_rles
        if len(target_instance_rles) > 0:               # labels continue after the largest one seen
            self.next_label = max(target_instance_rles.keys()) + 1

    def update_target(self, instance_rles):
        self.target_rle = instance_rles

    def _labels_for(self, incoming, hungarian, target_labels, ioa):
        """Rule 1-3 above for every incoming label; returns them in order."""
        by_hungarian = dict(zip(hungarian[1].tolist(), hungarian[0].tolist())) if len(hungarian[0]) else {}
        if len(ioa) > 0:
            best, holder = ioa.max(axis=0), ioa.argmax(axis=0)
        received = []
        for col, lab in enumerate(incoming):
            if lab in by_hungarian:
                received.append(by_hungarian[lab])
            elif len(ioa) > 0 and best[col] >= self.merge_ioa_thr:
                received.append(int(target_labels[holder[col]]))
            elif self.assign_new:
                received.append(self.next_label)
                self.next_label += 1
            else:
                received.append(lab)
        return received

    def __call__(self, match_instance_rle, update_target=True, inter=None, areas=None):
        assert self.target_rle is not None, "Initialize target rle before running!"
        hungarian, (target_labels, match_labels), _, ioa = rle_matcher(
            self.target_rle, match_instance_rle, self.merge_iou_thr, return_ioa=True, inter=inter, areas=areas)
        incoming = [int(k) for k in match_instance_rle.keys()]
        received = self._labels_for(incoming, hungarian, target_labels, ioa)
        self.last_assignment = list(zip(incoming, received))
        relabelled = {}
        for (old, new), attrs in zip(self.last_assignment, match_instance_rle.values()):
            relabelled[new] = merge_attrs(relabelled[new], attrs) if new in relabelled else attrs
        if update_target:
            self.update_target(relabelled)
        return relabelled


class StackMatcher:
    """Forward and backward matching of one class through a z-block (the loop of
    patterns.forward_matching / backward_matching, patterns.py:68-112) driven by precomputed overlaps
    of the ORIGINAL per-slice instances (``block_overlaps``): no run list is touched while matching.

    rles       list over slices of {label: attrs} for this class (dict order = instance slot order)
    overlaps   list over consecutive pairs of (slot_a, slot_b, inter) arrays
    """

    def __init__(self, class_id, label_divisor, merge_iou_thr=0.25, merge_ioa_thr=0.25):
        self.matcher = RLEMatcher(class_id, label_divisor, merge_iou_thr, merge_ioa_thr, True)

    # A slice's (possibly merged) instances are kept as `(labels, group_of_slot)`: the labels in dict
    # order and, for every ORIGINAL instance slot of the slice, the index of the label it now belongs to.
    @staticmethod
    def _group_inter(g_t, g_m, slot_t, slot_m, ov):
        """Non-zero (target label index, match label index, intersection) entries from slot-level overlaps:
        instances of a slice are disjoint, so a group's intersection is the sum over its slots."""
        n_m = max(len(g_m[0]), 1)
        key = g_t[1][slot_t] * n_m + g_m[1][slot_m]
        if key.size == 0:
            return key, key, ov
        uniq, inv = np.unique(key, return_inverse=True)
        return uniq // n_m, uniq % n_m, np.bincount(inv, weights=ov, minlength=uniq.size).astype(np.int64)

    @staticmethod
    def _group_areas(g, slot_areas):
        return np.bincount(g[1], weights=slot_areas, minlength=len(g[0])).astype(np.int64)

    @staticmethod
    def _regroup(g_old, assignment):
        """Groups after a matching step: `assignment` lists (old label, new label) in the old dict order."""
        new_labels, index = [], {}
        of_old = np.empty(len(assignment), np.int64)
        for i, (_, new) in enumerate(assignment):
            j = index.get(new)
            if j is None:
                j = index[new] = len(new_labels)
                new_labels.append(new)
            of_old[i] = j
        return new_labels, of_old[g_old[1]]

    def forward(self, rles, overlaps, slot_areas=None, prev=None):
        """Returns (matched rles per slice, groups per slice).
        slot_areas: optional list over slices of per-instance pixel counts (dict order).
        prev: the forward state of the block below (``forward_state`` of the previous rank) plus
        'overlaps' = (slot in its last slice, slot in rles[0], inter): the chain then continues from
        that block instead of starting at rles[0]."""
        m = self.matcher
        self.slot_areas = slot_areas if slot_areas is not None else [instance_areas(seg) for seg in rles]
        out, groups = [], []
        if prev is not None:
            m.target_rle, m.next_label = prev['target'], prev['next_label']
        for z, seg in enumerate(rles):
            own = (list(seg.keys()), np.arange(len(seg), dtype=np.int64))
            if m.target_rle is None:
                m.initialize_target(seg)
                out.append(seg)
                groups.append(own)
                continue
            if z == 0:
                g_t, a_t, (sa, sb, ov) = prev['groups'], prev['areas'], prev['overlaps']
            else:
                g_t, a_t, (sa, sb, ov) = groups[-1], self.slot_areas[z - 1], overlaps[z - 1]
            inter = self._group_inter(g_t, own, sa, sb, ov)
            areas = (self._group_areas(g_t, a_t), self.slot_areas[z])
            out.append(m(seg, inter=inter, areas=areas))
            groups.append(self._regroup(own, m.last_assignment))
        return out, groups

    def forward_state(self, groups):
        """What the next block needs to continue the forward chain (picklable)."""
        m = self.matcher
        return {'target': m.target_rle, 'next_label': m.next_label, 'groups': groups[-1], 'areas': self.slot_areas[-1]}

    def backward(self, fwd, groups, rles, overlaps, nxt=None):
        """patterns.backward_matching: targets reset, assign_new off, slices in reverse.
        nxt: the backward state of the block above (``backward_state`` of the next rank) plus
        'overlaps' = (slot in fwd[-1]'s slice, slot in its first slice, inter)."""
        m = self.matcher
        m.target_rle = None if nxt is None else nxt['target']
        m.assign_new = False
        n = len(fwd)
        out = [None] * n
        self.slot_labels = [None] * n                   # final label of every ORIGINAL instance slot, per slice
        g_next = None
        for z in range(n - 1, -1, -1):
            seg = fwd[z]
            if m.target_rle is None:
                m.initialize_target(seg)
                out[z], g_next = seg, groups[z]
            else:
                if z == n - 1:
                    g_t, a_t, (sa, sb, ov) = nxt['groups'], nxt['areas'], nxt['overlaps']
                else:
                    g_t, a_t, (sa, sb, ov) = g_next, self.slot_areas[z + 1], overlaps[z]
                inter = self._group_inter(g_t, groups[z], sb, sa, ov)      # pair (z, z+1): a = slots of z, b = slots of z+1
                areas = (self._group_areas(g_t, a_t), self._group_areas(groups[z], self.slot_areas[z]))
                out[z] = m(seg, inter=inter, areas=areas)
                g_next = self._regroup(groups[z], m.last_assignment)
            self.slot_labels[z] = np.asarray(g_next[0], np.int64)[g_next[1]] if len(g_next[1]) else np.zeros(0, np.int64)
        self._g_first = g_next
        return out

    def backward_state(self, out):
        """What the block below needs to continue the backward chain (picklable)."""
        return {'target': out[0], 'groups': self._g_first, 'areas': self.slot_areas[0]}
