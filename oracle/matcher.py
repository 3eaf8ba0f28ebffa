"""
oracle.matcher — CPU restatement of the reference's cross-slice RLE matcher
(empanada/inference/matcher.py:136-326, array_utils.py:101-125, :371-449, :625-718).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity: pinned — tests/golden/matcher_*.npz were
produced by the unmodified reference (tests/golden/make_golden.py matcher_cases, incl. the
reference's own tests/test_matcher.py known answer); tests/test_oracle_golden.py checks this module
against them.

Closed forms used instead of the reference's op sequence:
  * rle_intersection (array_utils.py:371-403, intersection_from_ranges :339-369): for two run lists
    that are each sorted and disjoint (what pan_seg_to_rle_seg and merge_rles produce) the sweep
    returns exactly |A n B| in pixels; computed here by clipping every run of B against A.
  * merge_rles (:690-718 -> join_ranges -> _join_ranges :634-663): union of the ranges, sorted by
    start, ranges that overlap or touch (end >= next start) joined.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment


def rle_intersection(starts_a, runs_a, starts_b, runs_b):
    sa, ea = np.asarray(starts_a, np.int64), np.asarray(starts_a, np.int64) + np.asarray(runs_a, np.int64)
    total = 0
    for s, r in zip(np.asarray(starts_b, np.int64), np.asarray(runs_b, np.int64)):
        ov = np.minimum(ea, s + r) - np.maximum(sa, s)
        total += int(ov[ov > 0].sum())
    return total


def merge_boxes(b1, b2):
    n = len(b1) // 2
    return tuple(min(b1[i], b2[i]) if i < n else max(b1[i], b2[i]) for i in range(len(b1)))


def merge_rles(starts_a, runs_a, starts_b, runs_b):
    s = np.concatenate([np.asarray(starts_a, np.int64), np.asarray(starts_b, np.int64)])
    e = s + np.concatenate([np.asarray(runs_a, np.int64), np.asarray(runs_b, np.int64)])
    order = np.argsort(s, kind='stable')
    s, e = s[order], e[order]
    out_s, out_e = [int(s[0])], [int(e[0])]
    for a, b in zip(s[1:], e[1:]):
        if out_e[-1] >= a:
            out_e[-1] = max(out_e[-1], int(b))
        else:
            out_s.append(int(a))
            out_e.append(int(b))
    out_s, out_e = np.asarray(out_s, np.int64), np.asarray(out_e, np.int64)
    return out_s, out_e - out_s


def merge_attrs(a1, a2):
    st, ru = merge_rles(a1['starts'], a1['runs'], a2['starts'], a2['runs'])
    return {'box': merge_boxes(a1['box'], a2['box']), 'starts': st, 'runs': ru}


def _boxes_overlap(b1, b2):
    n = len(b1) // 2
    return all(min(b1[i + n], b2[i + n]) - max(b1[i], b2[i]) > 0 for i in range(n))


def rle_matcher(target, match, iou_thr=0.5, return_iou=False, return_ioa=False):
    """matcher.py:136-232.  iou_matrix float64, ioa_matrix float32 (as the reference allocates them)."""
    tl = np.array([int(k) for k in target.keys()])
    ml = np.array([int(k) for k in match.keys()])
    if len(tl) == 0 or len(ml) == 0:
        empty = np.array([])
        if return_ioa:
            return (empty, empty), (tl, ml), empty, empty
        return (empty, empty), (tl, ml), empty
    ta, ma = list(target.values()), list(match.values())
    iou = np.zeros((len(tl), len(ml)), dtype='float')
    ioa = np.zeros((len(tl), len(ml)), dtype=np.float32)
    for i, a in enumerate(ta):
        for j, b in enumerate(ma):
            if not _boxes_overlap(a['box'], b['box']):       # box screening: box_iou(...).nonzero()
                continue
            inter = rle_intersection(a['starts'], a['runs'], b['starts'], b['runs'])
            iou[i, j] = inter / (int(np.sum(a['runs'])) + int(np.sum(b['runs'])) - inter)
            ioa[i, j] = inter / int(np.sum(b['runs']))
    rows, cols = linear_sum_assignment(iou, maximize=True)
    if iou_thr is not None:
        keep = iou[rows, cols] >= iou_thr
        rows, cols = rows[keep], cols[keep]
    out = ((tl[rows], ml[cols]), [tl, ml], iou[(rows, cols)])
    if return_iou:
        out = out + (iou,)
    if return_ioa:
        out = out + (ioa,)
    return out


class RLEMatcher:
    """matcher.py:234-326."""

    def __init__(self, class_id, label_divisor, merge_iou_thr=0.25, merge_ioa_thr=0.25, assign_new=True, **kwargs):
        self.class_id, self.label_divisor = class_id, label_divisor
        self.merge_iou_thr, self.merge_ioa_thr, self.assign_new = merge_iou_thr, merge_ioa_thr, assign_new
        self.next_label = class_id * label_divisor + 1
        self.target_rle = None

    def initialize_target(self, target):
        self.target_rle = target
        objs = list(target.keys())
        if len(objs) > 0:
            self.next_label = max(objs) + 1

    def update_target(self, rles):
        self.target_rle = rles

    def __call__(self, match, update_target=True):
        assert self.target_rle is not None, "Initialize target rle before running!"
        matched, all_labels, _, ioa = rle_matcher(self.target_rle, match, self.merge_iou_thr, return_ioa=True)
        tl, ml = all_labels
        by_match = {m: t for t, m in zip(matched[0], matched[1])}
        out = {}
        for i, (lab, attrs) in enumerate(match.items()):
            if lab in by_match:
                new = by_match[lab]
            else:
                ioa_max = ioa[:, i].max() if len(ioa) > 0 else 0
                if ioa_max >= self.merge_ioa_thr:
                    new = tl[ioa[:, i].argmax()]
                elif self.assign_new:
                    new = self.next_label
                    self.next_label += 1
                else:
                    new = lab
            out[new] = attrs if new not in out else merge_attrs(out[new], attrs)
        if update_target:
            self.update_target(out)
        return out
