"""
Drop-in for ``empanada.inference.filters`` (reference empanada/inference/filters.py:9-43): in-place
clean-up of a tracker's 3D instances before consensus / filling.  Host bookkeeping over a few thousand
dict entries; both filters decide on every instance first and delete afterwards.
"""
import numpy as np

__all__ = ['remove_small_objects', 'remove_pancakes']


def _drop(object_tracker, doomed):
    for instance_id in doomed:
        del object_tracker.instances[instance_id]


def remove_small_objects(object_tracker, min_size=64):
    """Delete instances with fewer than `min_size` voxels (sum of run lengths; filters.py:9-23)."""
    _drop(object_tracker, [i for i, attrs in object_tracker.instances.items()
                           if np.sum(attrs['runs']) < min_size])


def remove_pancakes(object_tracker, min_span=4):
    """Delete instances whose (z1, y1, x1, z2, y2, x2) box is thinner than `min_span` along any axis
    (filters.py:25-43)."""
    def thin(box):
        return min(box[3] - box[0], box[4] - box[1], box[5] - box[2]) < min_span

    _drop(object_tracker, [i for i, attrs in object_tracker.instances.items() if thin(attrs['box'])])
