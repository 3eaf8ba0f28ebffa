"""
Drop-in for ``empanada.inference.tracker`` (reference empanada/inference/tracker.py): collects the
per-slice run-length encoded instances of one class into 3D instances of a (d, h, w) volume, for stacks
taken along 'xy', 'xz' or 'yz', and (de)serialises them in the reference's json wire format.

This is host-side bookkeeping on a few thousand runs per slice (the GPU hands over RLEs, never dense
maps), written with array arithmetic instead of the reference's per-run Python loops.  Behaviours kept
exactly, because downstream consumers (consensus, filling) see them:
  * 'xy': a run keeps its length, its start moves by index2d * h * w (tracker.py:72-75);
  * 'xz': only the run START is lifted to 3D, the length is kept as is — a run that wraps around the
    end of a row of the (d, w) slice therefore does not wrap in the volume (tracker.py:76-81);
  * 'yz': runs are decoded to single voxels, lifted one by one, and re-encoded in ``finish`` after a
    stable sort, because consecutive voxels of a (d, h) slice are not consecutive in the volume
    (tracker.py:82-88, :108-113).
"""
import json
import math
from copy import deepcopy

import numpy as np

__all__ = ['InstanceTracker', 'to_box3d', 'rle_encode', 'rle_decode', 'rle_to_string', 'string_to_rle']

_AXES = {'xy': 0, 'xz': 1, 'yz': 2}


# ---- run-length helpers with the reference's semantics (array_utils.py:209-283) ---------------------
def rle_encode(indices):
    """Sorted 1d indices -> (starts, runs): a run breaks wherever idx[i] != idx[i-1] + 1."""
    indices = np.asarray(indices)
    if indices.size == 0:
        return indices[:0], np.zeros(0, dtype=np.int64)
    cut = np.flatnonzero(indices[1:] != indices[:-1] + 1) + 1
    first = np.concatenate(([0], cut))
    return indices[first], np.diff(np.concatenate((first, [indices.size])))


def rle_decode(starts, runs):
    """(starts, runs) -> every covered index, run after run."""
    starts, runs = np.asarray(starts, np.int64), np.asarray(runs, np.int64)
    if runs.size == 0:
        return np.zeros(0, np.int64)
    within = np.arange(int(runs.sum()), dtype=np.int64) - np.repeat(np.cumsum(runs) - runs, runs)
    return np.repeat(starts, runs) + within


def rle_to_string(starts, runs):
    """"s0 r0 s1 r1 ..." (array_utils.py:254-267)."""
    return ' '.join(f'{s} {r}' for s, r in zip(starts, runs))


def string_to_rle(encoding):
    flat = np.array([int(tok) for tok in encoding.split(' ')])
    return flat[::2], flat[1::2]


def to_box3d(index2d, box, axis):
    """A slice's (h1, w1, h2, w2) box as a one-voxel-thick 3D box at position index2d of `axis`."""
    assert axis in _AXES
    h1, w1, h2, w2 = box
    lo, hi = [h1, w1], [h2, w2]
    lo.insert(_AXES[axis], index2d)
    hi.insert(_AXES[axis], index2d + 1)
    return tuple(lo + hi)


def _merge_boxes(a, b):
    half = len(a) // 2
    return tuple(min(x, y) if i < half else max(x, y) for i, (x, y) in enumerate(zip(a, b)))


class InstanceTracker:
    def __init__(self, class_id=None, label_divisor=None, shape3d=None, axis='xy'):
        assert axis in ['xy', 'xz', 'yz']
        self.class_id = class_id
        self.label_divisor = label_divisor
        self.shape3d = shape3d
        self.axis = axis
        self.finished = False
        self.reset()
        self.axis_nums = {'xy': 0, 'xz': 1, 'yz': 2}

    def reset(self):
        self.instances = {}

    def _lift(self, starts, runs, index2d):
        """Flat indices of a 2D slice -> flat indices of the volume (see the module docstring)."""
        d, h, w = self.shape3d
        starts = np.asarray(starts)
        if self.axis == 'xy':
            return starts + index2d * (h * w), runs
        if self.axis == 'xz':                           # slice is (d, w); only the start is lifted
            return (starts // w) * (h * w) + index2d * w + starts % w, runs
        voxels = rle_decode(starts, runs)               # slice is (d, h): voxel by voxel
        lifted = (voxels // h) * (h * w) + (voxels % h) * w + index2d
        return lifted, np.ones_like(lifted)

    def update(self, instance_rles, index2d):
        assert self.class_id is not None
        assert self.label_divisor is not None
        assert self.shape3d is not None
        assert not self.finished, "Cannot update tracker after calling finish!"
        for label, attrs in instance_rles.items():
            box = to_box3d(index2d, attrs['box'], self.axis)
            starts, runs = self._lift(attrs['starts'], attrs['runs'], index2d)
            entry = self.instances.get(label)
            if entry is None:
                self.instances[label] = {'box': box, 'starts': [starts], 'runs': [runs]}
            else:
                entry['box'] = _merge_boxes(box, entry['box'])
                entry['starts'].append(starts)
                entry['runs'].append(runs)

    def finish(self):
        for entry in self.instances.values():
            if not isinstance(entry['starts'], list):
                continue                                # already concatenated
            starts = np.concatenate(entry['starts'])
            if self.axis == 'yz':                       # single voxels: sort and re-encode
                starts, runs = rle_encode(np.sort(starts, kind='stable'))
            else:
                runs = np.concatenate(entry['runs'])
            entry['starts'], entry['runs'] = starts, runs
        self.finished = True

    def write_to_json(self, savepath):
        if not self.finished:
            self.finish()
        save_dict = deepcopy(self.__dict__)
        packed = {}
        for label, entry in save_dict['instances'].items():
            entry['rle'] = rle_to_string(entry.pop('starts'), entry.pop('runs'))
            packed[str(label)] = entry
        save_dict['instances'] = packed
        with open(savepath, mode='w') as handle:
            json.dump(save_dict, handle, indent=6)

    def load_from_json(self, fpath):
        with open(fpath, mode='r') as handle:
            load_dict = json.load(handle)
        for entry in load_dict['instances'].values():
            entry['starts'], entry['runs'] = string_to_rle(entry['rle'])
        self.__dict__ = load_dict
