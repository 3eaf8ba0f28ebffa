"""
Drop-in for ``empanada.inference.patterns`` (reference empanada/inference/patterns.py): the glue the 3D
inference scripts (scripts/pdl_inference3d.py:120-330, pdl_multigpu_inference3d.py) call between the
engine and the output volumes — matchers and trackers per class, the forward / backward matching loops,
consensus trackers, volume filling, and the multi-GPU worker's post-processing loop.  Same names,
argument order and behaviour as the reference; what runs underneath is this package:

  pan_seg -> RLE              inference/rle.py       (emp_rle on the GPU)
  cross-slice matching        inference/matcher.py   (run-intersection kernel + host assignment)
  3D lifting / json           inference/tracker.py
  consensus                   consensus.py           (emp_rle_list_overlaps for the overlap graph)
  filling                     inference/fill.py      (emp_fill_runs)
  median + harden + merge     inference/engines.py   (emp_median_harden, emp_merge)

For whole z-blocks resident on one GPU the batched path is ``inference/stack.py`` (StackShard), which
gives the same results as these slice-at-a-time loops with one overlap launch per block.
"""
import numpy as np
import torch
import torch.distributed as dist

from empanada_b200.consensus import merge_objects_from_trackers, merge_semantic_from_trackers
from empanada_b200.inference import filters
from empanada_b200.inference.engines import _MedianQueue, median_harden
from empanada_b200.inference.fill import fill_instances, fill_slabs
from empanada_b200.inference.matcher import RLEMatcher
from empanada_b200.inference.postprocess import merge_semantic_and_instance
from empanada_b200.inference.rle import pan_seg_to_rle_seg, rle_seg_to_pan_seg  # noqa: F401
from empanada_b200.inference.tracker import InstanceTracker

__all__ = [
    'create_matchers',
    'create_axis_trackers',
    'apply_matchers',
    'forward_matching',
    'backward_matching',
    'update_trackers',
    'finish_tracking',
    'apply_filters',
    'get_axis_trackers_by_class',
    'create_instance_consensus',
    'create_semantic_consensus',
    'fill_volume',
    'fill_panoptic_volume',
    'all_gather',
    'forward_multigpu'
]


# ---- per-class matchers and trackers (patterns.py:33-66) ------------------------------------------------
def create_matchers(thing_list, label_divisor, merge_iou_thr, merge_ioa_thr):
    """One cross-slice matcher per thing class."""
    return [RLEMatcher(c, label_divisor, merge_iou_thr, merge_ioa_thr) for c in thing_list]


def create_axis_trackers(axes, class_labels, label_divisor, shape):
    """{axis name ('xy' | 'xz' | 'yz'): [one tracker per class]}."""
    return {name: [InstanceTracker(c, label_divisor, shape, name) for c in class_labels] for name in axes}


def apply_matchers(rle_seg, matchers):
    """Relabel every thing class of one slice against the slice its matcher saw last; a matcher without a
    target takes this slice as its first target and leaves it untouched."""
    for m in matchers:
        if m.target_rle is None:
            m.initialize_target(rle_seg[m.class_id])
        else:
            rle_seg[m.class_id] = m(rle_seg[m.class_id])
    return rle_seg


# ---- matching loops (patterns.py:68-134) ---------------------------------------------------------------
def _encode_and_match(pan_seg, matchers, labels, label_divisor, thing_list):
    return apply_matchers(pan_seg_to_rle_seg(pan_seg, labels, label_divisor, thing_list, force_connected=True), matchers)


def forward_matching(matchers, queue, rle_stack, matcher_in, labels, label_divisor, thing_list):
    """Consumer loop of the matching process: panoptic maps arrive on `queue` (None while the engine's
    median window is filling, any string to stop), each is encoded and matched to its predecessor; the
    finished stack goes back through the pipe end `matcher_in`."""
    while True:
        pan_seg = queue.get()
        if pan_seg is None:
            continue
        if isinstance(pan_seg, str):
            break
        rle_stack.append(_encode_and_match(pan_seg, matchers, labels, label_divisor, thing_list))
    matcher_in.send([rle_stack])
    matcher_in.close()


def backward_matching(rle_stack, matchers, axis_len):
    """Generator over (index, rle_seg) from the last slice to the first: labels propagate backward with no
    new labels handed out."""
    for m in matchers:
        m.target_rle = None
        m.assign_new = False
    for index in range(axis_len - 1, -1, -1):
        yield index, apply_matchers(rle_stack[index], matchers)


def update_trackers(rle_seg, index, trackers, *unused):
    """Hand slice `index`'s matched instances to the tracker of each class.  (The reference script passes
    two more positional arguments than its own function takes, scripts/pdl_inference3d.py:196; they are
    accepted and ignored here.)"""
    for tr in trackers:
        tr.update(rle_seg[tr.class_id], index)


def finish_tracking(trackers):
    for tr in trackers:
        tr.finish()


def apply_filters(tracker, filters_dict):
    """filters_dict: list of {'name': <function in inference.filters>, **kwargs}; applied in place."""
    for spec in filters_dict or ():
        kwargs = {k: v for k, v in spec.items() if k != 'name'}
        getattr(filters, spec['name'])(tracker, **kwargs)


# ---- consensus trackers (patterns.py:154-202) ---------------------------------------------------------
def get_axis_trackers_by_class(trackers, class_id):
    """The trackers of one class across all axes, in axis order."""
    return [tr for axis_trackers in trackers.values() for tr in axis_trackers if tr.class_id == class_id]


def _like_first(class_trackers):
    first = class_trackers[0]
    return InstanceTracker(first.class_id, first.label_divisor, first.shape3d, 'xy')


def create_instance_consensus(class_trackers, pixel_vote_thr=2, cluster_iou_thr=0.75, bypass=False):
    out = _like_first(class_trackers)
    out.instances = merge_objects_from_trackers(class_trackers, pixel_vote_thr, cluster_iou_thr, bypass)
    return out


def create_semantic_consensus(class_trackers, pixel_vote_thr=2):
    out = _like_first(class_trackers)
    out.instances = merge_semantic_from_trackers(class_trackers, pixel_vote_thr)
    return out


# ---- filling (patterns.py:204-222) ---------------------------------------------------------------------
def fill_volume(volume, instances, processes=4):
    """Paint run-length encoded instances into `volume` in place (patterns.py:160-172 -> array_utils.numpy_fill_instances /
    zarr_utils.zarr_fill_instances).  A CUDA tensor is painted where it is.  A numpy array or a zarr array (anything
    with .shape / .dtype that reads and writes z-slabs by slicing) is filled one z-slab at a time through HBM, whole
    z-chunks per slab for chunked stores, so the volume never has to fit into GPU memory.  `processes` (the
    reference's pool size for zarr stores) is accepted and unused: the painting is one kernel launch per slab."""
    if torch.is_tensor(volume):
        fill_instances(volume, instances)
    elif isinstance(volume, np.ndarray) or (hasattr(volume, 'shape') and hasattr(volume, 'dtype') and hasattr(volume, '__setitem__')):
        if len(instances) == 0:
            return
        if len(volume.shape) != 3:
            raise Exception(f'Expected a (d, h, w) volume, got shape {tuple(volume.shape)}')
        fill_slabs(volume, instances)
    else:
        raise Exception(f'Unknown volume type of {type(volume)}')


def fill_panoptic_volume(volume, trackers, processes=4):
    for tr in trackers:
        fill_volume(volume, tr.instances, processes)


# ---- multi-GPU worker (patterns.py:224-350) ------------------------------------------------------------
def all_gather(tensor, group=None):
    """List of every rank's `tensor` (just [tensor]-shaped when no process group is up)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    gathered = [torch.zeros_like(tensor) for _ in range(world)]
    if world == 1:
        gathered[0].copy_(tensor)
    else:
        dist.all_gather(gathered, tensor, group=group)
    return gathered


def harden_seg(sem, confidence_thr):
    """(1,C,H,W) probabilities -> (1,1,H,W) int64: first-max argmax for C > 1, `>= thr` for C == 1."""
    return median_harden([sem], confidence_thr, want_median=False, want_sem='i64')[1]


def get_panoptic_seg(sem, instance_cells, label_divisor, thing_list, stuff_area=32, void_label=0):
    """Hardened sem + float instance cells -> panoptic map (patterns.py:253-277)."""
    thing = torch.zeros_like(sem)
    for c in thing_list:
        thing[sem == c] = 1
    return merge_semantic_and_instance(sem, (thing * instance_cells).long(), label_divisor, thing_list,
                                       stuff_area, void_label)


def forward_multigpu(matchers, queue, rle_stack, matcher_in, confidence_thr, median_kernel_size, labels,
                     label_divisor, thing_list, stuff_area=32, void_label=0):
    """Worker loop of the multi-GPU script: (sem probabilities, instance cells) pairs arrive in z order from
    the gathering rank; runs the median window, hardening, merge, RLE encoding and forward matching."""
    window = _MedianQueue(median_kernel_size)

    def on_gpu(t):
        return t if t.is_cuda else t.cuda()

    def emit(entry):
        sem = harden_seg(entry['sem'], confidence_thr)
        pan_seg = get_panoptic_seg(sem, on_gpu(entry['cells']), label_divisor, thing_list, stuff_area, void_label)
        rle_stack.append(_encode_and_match(pan_seg.squeeze(), matchers, labels, label_divisor, thing_list))

    while True:
        sem, cells = queue.get()
        if isinstance(sem, str):
            break
        window.enqueue({'sem': on_gpu(sem), 'cells': cells})
        entry = window.get_next(keys=['sem'])
        if entry is not None:
            emit(entry)
    for entry in window.end():
        emit(entry)
    matcher_in.send([rle_stack])
    matcher_in.close()
