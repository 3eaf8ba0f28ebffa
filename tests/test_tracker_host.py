"""Host logic of the tracker drop-in against a fixture produced by the reference's own
InstanceTracker (tests/golden/tracker_axes.npz): 3D lifting along xy / xz / yz incl. the xz row-wrap
behaviour and the yz re-encode, and the json wire format byte for byte."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import load_golden
from empanada_b200.inference import tracker as tk


@pytest.mark.parametrize('axis', ['xy', 'xz', 'yz'])
def test_tracker_matches_reference(axis, tmp_path):
    g = load_golden('tracker_axes')
    vol = g['in_vol']
    tr = tk.InstanceTracker(1, 1000, vol.shape, axis=axis)
    n = vol.shape[{'xy': 0, 'xz': 1, 'yz': 2}[axis]]
    for i in range(n):
        sl = vol[i] if axis == 'xy' else vol[:, i] if axis == 'xz' else vol[..., i]
        tr.update(oracle.pan_seg_to_rle_seg(np.ascontiguousarray(sl), [1], 1000, [1], False)[1], i)
    with pytest.raises(AssertionError):
        tk.InstanceTracker(1, 1000, vol.shape, axis='zz')
    tr.finish()
    with pytest.raises(AssertionError):
        tr.update({}, 0)
    labs = list(tr.instances.keys())
    np.testing.assert_array_equal(np.asarray(labs, np.int64), g[f'{axis}_labels'])
    np.testing.assert_array_equal(np.asarray([tr.instances[l]['box'] for l in labs], np.int64), g[f'{axis}_boxes'])
    np.testing.assert_array_equal(np.asarray([len(tr.instances[l]['starts']) for l in labs]), g[f'{axis}_counts'])
    np.testing.assert_array_equal(np.concatenate([tr.instances[l]['starts'] for l in labs]), g[f'{axis}_starts'])
    np.testing.assert_array_equal(np.concatenate([tr.instances[l]['runs'] for l in labs]), g[f'{axis}_runs'])
    if axis == 'xy':
        path = os.path.join(str(tmp_path), 't.json')
        tr.write_to_json(path)
        assert open(path, 'rb').read() == g['xy_json'].tobytes()
        back = tk.InstanceTracker()
        back.load_from_json(path)
        assert back.axis == 'xy' and back.finished and list(back.shape3d) == list(vol.shape)
        for l in labs:
            np.testing.assert_array_equal(back.instances[str(l)]['starts'], tr.instances[l]['starts'])
            np.testing.assert_array_equal(back.instances[str(l)]['runs'], tr.instances[l]['runs'])
        # filling the tracked instances reproduces the volume (tests/test_tracking.py of the reference)
        out = np.zeros(vol.size, np.int64)
        for l in labs:
            for s, r in zip(tr.instances[l]['starts'], tr.instances[l]['runs']):
                out[s:s + r] = l
        np.testing.assert_array_equal(out.reshape(vol.shape), vol)


def test_rle_helpers():
    idx = np.array([3, 4, 5, 9, 10, 20])
    s, r = tk.rle_encode(idx)
    np.testing.assert_array_equal(s, [3, 9, 20])
    np.testing.assert_array_equal(r, [3, 2, 1])
    np.testing.assert_array_equal(tk.rle_decode(s, r), idx)
    assert tk.rle_to_string(s, r) == '3 3 9 2 20 1'
    s2, r2 = tk.string_to_rle('3 3 9 2 20 1')
    np.testing.assert_array_equal(s2, s)
    np.testing.assert_array_equal(r2, r)
    assert tk.to_box3d(7, (1, 2, 3, 4), 'xz') == (1, 7, 2, 3, 8, 4)
