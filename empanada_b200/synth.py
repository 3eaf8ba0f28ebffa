"""
Synthetic EM-shaped head tensors (numpy only).

Random-init PanopticDeepLab heads emit constants (the final 1x1 convs are ~N(0, 0.001^2),
reference models/heads.py:21-26), so a benchmark that wants post-processing *work* needs head
tensors with real structure.  This generator mirrors what the reference trains against
(data/utils/target_creation.py:13-78): an instance map of filled ellipses, a center heat-map
(one Gaussian of sigma 6 per instance centroid, normalised to max 1) and per-pixel offsets to
the instance centroid (0 on background), plus N(0, noise) jitter on the offsets so the
nearest-center search is not trivially exact.

All randomness is ``numpy.random.default_rng(seed)``; the same seed gives the same tile.
"""
import numpy as np

__all__ = ['synth_instances', 'synth_tile', 'synth_stack_slices', 'synth_heads', 'synth_blob_volume']


def synth_instances(H, W, n_instances, rng, semi_axes=(12.0, 40.0)):
    """(H,W) int32 instance map: n filled ellipses, later ones overwrite earlier ones."""
    ins = np.zeros((H, W), np.int32)
    cy = rng.uniform(0, H, n_instances)
    cx = rng.uniform(0, W, n_instances)
    a = rng.uniform(semi_axes[0], semi_axes[1], n_instances)
    b = rng.uniform(semi_axes[0], semi_axes[1], n_instances)
    th = rng.uniform(0, np.pi, n_instances)
    for i in range(n_instances):
        r = int(np.ceil(max(a[i], b[i]))) + 1
        y0, y1 = max(0, int(cy[i]) - r), min(H, int(cy[i]) + r + 1)
        x0, x1 = max(0, int(cx[i]) - r), min(W, int(cx[i]) + r + 1)
        if y0 >= y1 or x0 >= x1:
            continue
        yy = np.arange(y0, y1, dtype=np.float32)[:, None] - np.float32(cy[i])
        xx = np.arange(x0, x1, dtype=np.float32)[None, :] - np.float32(cx[i])
        c, s = np.float32(np.cos(th[i])), np.float32(np.sin(th[i]))
        u = (xx * c + yy * s) / np.float32(a[i])
        v = (-xx * s + yy * c) / np.float32(b[i])
        ins[y0:y1, x0:x1][u * u + v * v <= 1.0] = i + 1
    return ins


def _centroids(ins):
    n = int(ins.max())
    flat = ins.ravel()
    H, W = ins.shape
    cnt = np.bincount(flat, minlength=n + 1).astype(np.float64)
    ys = np.repeat(np.arange(H, dtype=np.float64), W)
    xs = np.tile(np.arange(W, dtype=np.float64), H)
    sy = np.bincount(flat, weights=ys, minlength=n + 1)
    sx = np.bincount(flat, weights=xs, minlength=n + 1)
    ok = cnt > 0
    ok[0] = False
    cyc = np.zeros(n + 1)
    cxc = np.zeros(n + 1)
    cyc[ok] = sy[ok] / cnt[ok]
    cxc[ok] = sx[ok] / cnt[ok]
    return ok, cyc, cxc


def synth_tile(H, W, n_instances, seed, semi_axes=(12.0, 40.0), sigma=6.0, noise=0.5,
               thing_classes=(1,), stuff_classes=(), prob=False):
    """One tile of synthetic head tensors.

    Returns a dict of numpy arrays:
      'sem'     (1,1,H,W) int64 hardened semantic classes (0 = background),
      'ctr_hmp' (1,1,H,W) float32, 'offsets' (1,2,H,W) float32 (ch0 = dy, ch1 = dx),
      'ins'     (H,W) int32 ground-truth instance map,
      'sem_prob' (1,1,H,W) float32 — only if ``prob`` and a single thing class: 0.9 inside /
                 0.1 outside + U(-0.05, 0.05), for the median / confidence-threshold path.
    """
    rng = np.random.default_rng(seed)
    ins = synth_instances(H, W, n_instances, rng, semi_axes)
    ok, cyc, cxc = _centroids(ins)

    # heat-map: sum of truncated Gaussians at the int-truncated centroids, normalised to max 1
    hm = np.zeros((H, W), np.float32)
    r = int(np.ceil(4 * sigma))
    g1 = np.exp(-(np.arange(-r, r + 1, dtype=np.float32) ** 2) / np.float32(2 * sigma * sigma))
    g2 = np.outer(g1, g1).astype(np.float32)
    for i in np.flatnonzero(ok):
        y, x = int(cyc[i]), int(cxc[i])
        y0, y1 = max(0, y - r), min(H, y + r + 1)
        x0, x1 = max(0, x - r), min(W, x + r + 1)
        hm[y0:y1, x0:x1] += g2[y0 - (y - r):y1 - (y - r), x0 - (x - r):x1 - (x - r)]
    m = hm.max()
    if m > 0:
        hm /= m

    # offsets: centroid - pixel on things (+ jitter), 0 on background
    thing = ins > 0
    off = np.zeros((2, H, W), np.float32)
    yy, xx = np.nonzero(thing)
    lab = ins[yy, xx]
    off[0, yy, xx] = (cyc[lab] - yy).astype(np.float32)
    off[1, yy, xx] = (cxc[lab] - xx).astype(np.float32)
    if noise > 0:
        off[0, yy, xx] += rng.standard_normal(yy.size, dtype=np.float32) * np.float32(noise)
        off[1, yy, xx] += rng.standard_normal(yy.size, dtype=np.float32) * np.float32(noise)

    # semantic classes: each instance takes one thing class; stuff classes fill random boxes
    sem = np.zeros((H, W), np.int64)
    tc = np.asarray(thing_classes, np.int64)
    cls_of = np.zeros(int(ins.max()) + 1, np.int64)
    cls_of[1:] = tc[rng.integers(0, tc.size, cls_of.size - 1)]
    for s in stuff_classes:
        for _ in range(3):
            y0, x0 = int(rng.integers(0, H)), int(rng.integers(0, W))
            h, w = int(rng.integers(2, max(3, H // 3))), int(rng.integers(2, max(3, W // 3)))
            sem[y0:y0 + h, x0:x0 + w] = s
    sem[thing] = cls_of[ins[thing]]

    out = {'sem': sem[None, None], 'ctr_hmp': hm[None, None], 'offsets': off[None], 'ins': ins}
    if prob:
        p = np.where(thing, np.float32(0.9), np.float32(0.1)).astype(np.float32)
        p += rng.uniform(-0.05, 0.05, p.shape).astype(np.float32)
        out['sem_prob'] = p[None, None]
    return out


def synth_stack_slices(D, H, W, n_blobs, seed, coarse=4, sigma=6.0, noise=0.5,
                       z_extent=(8, 40), semi_axes=(12.0, 40.0)):
    """Generator over D slices of a synthetic anisotropic volume of ellipsoidal blobs, shaped
    like what the Render engines consume (reference quantization/panoptic_deeplab.py:221-250):
    yields dicts with 'sem_prob' (1,1,H,W) fp32 at full resolution and 'ctr_hmp' (1,1,h,w),
    'offsets' (1,2,h,w) at 1/``coarse`` resolution with offsets in *full-res* pixel units
    (the coarse grid is scaled by step=4 in group_pixels, reference engines.py:263-271).
    """
    rng = np.random.default_rng(seed)
    h, w = H // coarse, W // coarse
    zc = rng.uniform(0, D, n_blobs)
    zr = rng.uniform(z_extent[0], z_extent[1], n_blobs) / 2
    cy = rng.uniform(0, H, n_blobs)
    cx = rng.uniform(0, W, n_blobs)
    a = rng.uniform(semi_axes[0], semi_axes[1], n_blobs)
    b = rng.uniform(semi_axes[0], semi_axes[1], n_blobs)
    r = int(np.ceil(4 * sigma / coarse))
    sg = sigma / coarse
    g1 = np.exp(-(np.arange(-r, r + 1, dtype=np.float32) ** 2) / np.float32(2 * sg * sg))
    g2 = np.outer(g1, g1).astype(np.float32)
    for z in range(D):
        srng = np.random.default_rng([seed, z])
        ins = np.zeros((H, W), np.int32)
        live = np.flatnonzero(np.abs(z - zc) < zr)
        for i in live:
            f = np.sqrt(max(0.0, 1.0 - ((z - zc[i]) / zr[i]) ** 2))
            ai, bi = max(1.0, a[i] * f), max(1.0, b[i] * f)
            rr = int(np.ceil(max(ai, bi))) + 1
            y0, y1 = max(0, int(cy[i]) - rr), min(H, int(cy[i]) + rr + 1)
            x0, x1 = max(0, int(cx[i]) - rr), min(W, int(cx[i]) + rr + 1)
            if y0 >= y1 or x0 >= x1:
                continue
            yy = (np.arange(y0, y1, dtype=np.float32)[:, None] - np.float32(cy[i])) / np.float32(ai)
            xx = (np.arange(x0, x1, dtype=np.float32)[None, :] - np.float32(cx[i])) / np.float32(bi)
            ins[y0:y1, x0:x1][yy * yy + xx * xx <= 1.0] = i + 1
        ok, cyc, cxc = _centroids(ins) if ins.max() > 0 else (np.zeros(1, bool), np.zeros(1), np.zeros(1))
        hm = np.zeros((h, w), np.float32)
        for i in np.flatnonzero(ok):
            y, x = int(cyc[i]) // coarse, int(cxc[i]) // coarse
            y0, y1 = max(0, y - r), min(h, y + r + 1)
            x0, x1 = max(0, x - r), min(w, x + r + 1)
            hm[y0:y1, x0:x1] += g2[y0 - (y - r):y1 - (y - r), x0 - (x - r):x1 - (x - r)]
        m = hm.max()
        if m > 0:
            hm /= m
        # coarse offsets: from the coarse pixel's full-res coordinate (4*y, 4*x) to a centre
        # snapped to the coarse lattice the heat-map peaks live on
        cins = ins[::coarse, ::coarse]
        off = np.zeros((2, h, w), np.float32)
        yy, xx = np.nonzero(cins)
        lab = cins[yy, xx]
        ty = (np.floor(cyc[lab]) // coarse) * coarse
        tx = (np.floor(cxc[lab]) // coarse) * coarse
        off[0, yy, xx] = (ty - yy * coarse).astype(np.float32)
        off[1, yy, xx] = (tx - xx * coarse).astype(np.float32)
        if noise > 0:
            off[0, yy, xx] += srng.standard_normal(yy.size, dtype=np.float32) * np.float32(noise)
            off[1, yy, xx] += srng.standard_normal(yy.size, dtype=np.float32) * np.float32(noise)
        p = np.where(ins > 0, np.float32(0.9), np.float32(0.1)).astype(np.float32)
        p += srng.uniform(-0.05, 0.05, p.shape).astype(np.float32)
        yield {'sem_prob': p[None, None], 'ctr_hmp': hm[None, None], 'offsets': off[None], 'ins': ins}


def synth_heads(ins, rng, sigma=2.5, noise=0.5, prob_noise=0.22, prob_step=1.0 / 64, off_step=1.0 / 8):
    """Head tensors for ONE given (H,W) instance slice (any axis of a volume): 'sem_prob' (1,1,H,W),
    'ctr_hmp' (1,1,H,W), 'offsets' (1,2,H,W), all float32.  The jitter is quantised (prob_step / off_step)
    so that fixtures holding many slices compress well; ties and near-ties become MORE likely that way."""
    H, W = ins.shape
    hm = np.zeros((H, W), np.float32)
    off = np.zeros((2, H, W), np.float32)
    if ins.max() > 0:
        ok, cyc, cxc = _centroids(ins)
        r = int(np.ceil(4 * sigma))
        g1 = np.exp(-(np.arange(-r, r + 1, dtype=np.float32) ** 2) / np.float32(2 * sigma * sigma))
        g2 = np.outer(g1, g1).astype(np.float32)
        for i in np.flatnonzero(ok):
            y, x = int(cyc[i]), int(cxc[i])
            y0, y1 = max(0, y - r), min(H, y + r + 1)
            x0, x1 = max(0, x - r), min(W, x + r + 1)
            hm[y0:y1, x0:x1] += g2[y0 - (y - r):y1 - (y - r), x0 - (x - r):x1 - (x - r)]
        hm /= hm.max()
        yy, xx = np.nonzero(ins)
        lab = ins[yy, xx]
        jit = rng.standard_normal((2, yy.size)) * noise
        off[0, yy, xx] = np.round((cyc[lab] - yy + jit[0]) / off_step) * off_step
        off[1, yy, xx] = np.round((cxc[lab] - xx + jit[1]) / off_step) * off_step
    p = np.where(ins > 0, 0.9, 0.1) + rng.uniform(-prob_noise, prob_noise, ins.shape)
    p = np.clip(np.round(p / prob_step) * prob_step, prob_step, 1 - prob_step).astype(np.float32)
    return {'sem_prob': p[None, None], 'ctr_hmp': hm[None, None], 'offsets': off[None]}


def synth_blob_volume(shape, n_blobs, seed, radii=(4.0, 9.0)):
    """(D,H,W) int32 instance volume of axis-aligned ellipsoids (later ones overwrite earlier ones)."""
    rng = np.random.default_rng(seed)
    D, H, W = shape
    zz, yy, xx = np.mgrid[0:D, 0:H, 0:W].astype(np.float32)
    vol = np.zeros(shape, np.int32)
    c = rng.uniform(0, 1, (n_blobs, 3)) * np.array(shape)
    rad = rng.uniform(radii[0], radii[1], (n_blobs, 3))
    for i in range(n_blobs):
        m = ((zz - c[i, 0]) / rad[i, 0]) ** 2 + ((yy - c[i, 1]) / rad[i, 1]) ** 2 + ((xx - c[i, 2]) / rad[i, 2]) ** 2 <= 1
        vol[m] = i + 1
    return vol


# BASELINE configs[0] ("config 1"): the reference's own CPU-runnable case — one 1024 x 1024 tile, PanopticDeepLabEngine
# defaults (engines.py:92-112), ~60 instances.  tests/golden/make_config1.py runs the reference on it.
CONFIG1 = dict(thing_list=[1], label_divisor=1000, stuff_area=64, void_label=0, nms_threshold=0.1, nms_kernel=7, confidence_thr=0.5)
CONFIG1_TILE = (1024, 60, 0)            # side, instances, seed


def config1_heads(consts):
    """Config 1's head tensors as float32 numpy arrays: the random-initialised network's per-channel constants
    (consts: sem logit, heat-map, dy, dx — stored in tests/golden/config1.npz) + seeded synthetic heads.  No
    transcendental function on the way, so the values are the same bits on any host."""
    hw, n, seed = CONFIG1_TILE
    d = synth_tile(hw, hw, n, seed=seed)
    rng = np.random.default_rng(seed + 1)
    logit = np.where(d['ins'] > 0, np.float32(2.0), np.float32(-2.0)).astype(np.float32)
    logit += rng.uniform(-0.5, 0.5, logit.shape).astype(np.float32)
    c = np.asarray(consts, np.float32)
    return {'sem_logits': (logit + c[0])[None, None], 'ctr_hmp': d['ctr_hmp'] + c[1],
            'offsets': d['offsets'] + c[2:4].reshape(1, 2, 1, 1)}
