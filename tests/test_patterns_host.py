"""Host-only pieces of the patterns / filters drop-ins (no GPU): tracker bookkeeping helpers and the two
in-place filters (reference empanada/inference/filters.py:9-43, patterns.py:44-66,136-166)."""
import numpy as np

from empanada_b200.inference import filters, patterns


def _tracker_with(instances):
    tr = patterns.create_axis_trackers({'xy': 0}, [1], 1000, (8, 8, 8))['xy'][0]
    tr.instances = instances
    return tr


def test_create_axis_trackers_and_lookup():
    trackers = patterns.create_axis_trackers({'xy': 0, 'xz': 1, 'yz': 2}, [1, 2], 1000, (4, 5, 6))
    assert list(trackers) == ['xy', 'xz', 'yz']
    assert [t.class_id for t in trackers['xz']] == [1, 2]
    assert all(t.axis == name and t.shape3d == (4, 5, 6) for name, ts in trackers.items() for t in ts)
    twos = patterns.get_axis_trackers_by_class(trackers, 2)
    assert [t.axis for t in twos] == ['xy', 'xz', 'yz'] and all(t.class_id == 2 for t in twos)


def test_remove_small_objects_is_strict_less_than():
    inst = {1001: {'box': (0, 0, 0, 4, 4, 4), 'starts': np.array([0, 10]), 'runs': np.array([30, 34])},   # 64 voxels
            1002: {'box': (0, 0, 0, 4, 4, 4), 'starts': np.array([100]), 'runs': np.array([63])},
            1003: {'box': (0, 0, 0, 4, 4, 4), 'starts': np.array([200]), 'runs': np.array([65])}}
    tr = _tracker_with(dict(inst))
    filters.remove_small_objects(tr, min_size=64)
    assert list(tr.instances) == [1001, 1003]


def test_remove_pancakes_any_axis():
    def attrs(box):
        return {'box': box, 'starts': np.array([0]), 'runs': np.array([100])}
    tr = _tracker_with({1: attrs((0, 0, 0, 4, 4, 4)), 2: attrs((0, 0, 0, 3, 9, 9)), 3: attrs((0, 0, 0, 9, 3, 9)),
                        4: attrs((0, 0, 0, 9, 9, 3)), 5: attrs((2, 2, 2, 6, 7, 8))})
    filters.remove_pancakes(tr, min_span=4)
    assert list(tr.instances) == [1, 5]


def test_apply_filters_by_name():
    tr = _tracker_with({1: {'box': (0, 0, 0, 9, 9, 9), 'starts': np.array([0]), 'runs': np.array([5])},
                        2: {'box': (0, 0, 0, 9, 9, 1), 'starts': np.array([9]), 'runs': np.array([500])},
                        3: {'box': (0, 0, 0, 9, 9, 9), 'starts': np.array([9]), 'runs': np.array([500])}})
    patterns.apply_filters(tr, [{'name': 'remove_small_objects', 'min_size': 10}, {'name': 'remove_pancakes', 'min_span': 2}])
    assert list(tr.instances) == [3]
    patterns.apply_filters(tr, None)
    assert list(tr.instances) == [3]


def test_update_and_finish_trackers_accept_script_call_shape():
    trackers = patterns.create_axis_trackers({'xy': 0}, [1], 1000, (3, 4, 5))['xy']
    seg = {1: {1001: {'box': (0, 1, 2, 4), 'starts': np.array([1, 6]), 'runs': np.array([3, 3])}}}
    patterns.update_trackers(seg, 2, trackers, 0, None)          # the script passes (…, axis, stack) too
    patterns.finish_tracking(trackers)
    got = trackers[0].instances[1001]
    assert tuple(got['box']) == (2, 0, 1, 3, 2, 4)
    np.testing.assert_array_equal(got['starts'], np.array([41, 46]))
    np.testing.assert_array_equal(got['runs'], np.array([3, 3]))
