"""
ref_shim — lets the UNMODIFIED reference modules run in the build container.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py, never on the GPU box and never
by the product).  The reference's ``rle.py`` / ``matcher.py`` / ``patterns.py`` /
``data/utils/target_creation.py`` import scikit-image, cc3d and zarr, none of which are
installed here and there is no network.  This module injects just enough of
``skimage.measure`` (``label``, ``regionprops``), ``skimage.morphology.dilation`` and an empty
``zarr`` into ``sys.modules`` so those files import and run as they are:

  * ``label``        — full-connectivity multi-value CCL, numbered 1..n by the raster order of
                       each component's first pixel (what skimage's and cc3d's two-pass
                       union-find + sequential renumber produce), built on scipy.ndimage.label.
  * ``regionprops``  — objects with ``.label``, ``.bbox``, ``.coords`` (row-major) and
                       ``.centroid``, in ascending label order, label 0 skipped.

Because CCL numbering under ``force_connected=True`` is exercised by no reference test, the
contract "raster-first numbering" is the documented behaviour of both libraries rather than
something the reference pins; DESIGN.md says so.
"""
import sys
import types

import numpy as np
from scipy import ndimage

REFERENCE_ROOT = "/root/reference"


class _RegionProp:
    def __init__(self, lab, coords):
        self.label = int(lab)
        self.coords = coords
        self.bbox = tuple(int(v) for v in coords.min(0)) + tuple(int(v) + 1 for v in coords.max(0))
        self.centroid = tuple(float(v) for v in coords.mean(0))
        self.area = coords.shape[0]


def regionprops(img):
    img = np.asarray(img)
    flat = img.ravel()
    idx = np.flatnonzero(flat)
    if idx.size == 0:
        return []
    idx = idx[np.argsort(flat[idx], kind='stable')]
    labs = flat[idx]
    cuts = np.flatnonzero(np.diff(labs)) + 1
    return [_RegionProp(flat[g[0]], np.stack(np.unravel_index(g, img.shape), 1))
            for g in np.split(idx, cuts)]


def label(seg, **kwargs):
    seg = np.asarray(seg)
    st = np.ones((3,) * seg.ndim, bool)
    comps = []
    for v in np.unique(seg[seg != 0]):
        lab, n = ndimage.label(seg == v, structure=st)
        fl = lab.ravel()
        pos = np.flatnonzero(fl)
        first = np.full(n + 1, fl.size, np.int64)
        np.minimum.at(first, fl[pos], pos)
        comps += [(first[c], lab, c) for c in range(1, n + 1)]
    out = np.zeros(seg.shape, np.int64)
    for new, (_, lab, c) in enumerate(sorted(comps, key=lambda t: t[0]), 1):
        out[lab == c] = new
    return out


def install():
    """Insert the shim modules and put the reference on sys.path.  Idempotent."""
    if 'skimage' not in sys.modules:
        sk = types.ModuleType('skimage')
        me = types.ModuleType('skimage.measure')
        me.label = label
        me.regionprops = regionprops
        mo = types.ModuleType('skimage.morphology')
        mo.dilation = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError('shim'))
        io = types.ModuleType('skimage.io')
        sk.measure, sk.morphology, sk.io = me, mo, io
        sys.modules.update({'skimage': sk, 'skimage.measure': me,
                            'skimage.morphology': mo, 'skimage.io': io})
    if 'zarr' not in sys.modules:
        z = types.ModuleType('zarr')
        z.Array = type('Array', (), {})
        sys.modules['zarr'] = z
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
