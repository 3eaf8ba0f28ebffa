"""
torch_port — the reference's CPU *op sequence* for panoptic post-processing, restated with
torch CPU ops.  TEST INFRASTRUCTURE / BASELINE ONLY (see oracle/__init__.py).

Why it exists: bench.py must time "the reference's own CPU implementation of the path" beside
the GPU arm, but /root/reference does not travel to the GPU box.  oracle.c is a closed-form
restatement (one argmin per pixel) and would flatter the reference by orders of magnitude, so
this module restates what the reference actually executes — the same torch operators in the
same order with the same temporaries:

  * centers    threshold -> max_pool2d -> equality mask -> nonzero        (postprocess.py:55-75)
  * grouping   coordinate grid + offsets, then per chunk of 20 centers a (20, H*W, 2)
               difference tensor -> norm -> min -> masked update            (postprocess.py:97-116, :146-167)
  * merge      unique ids -> per-instance mask / count / mode / paste, then per stuff class
               mask / area / paste                                          (postprocess.py:253-294)

It is validated against the reference-generated goldens in tests/test_oracle_golden.py
(test_torch_port_matches_reference), so it is a faithful "port", not a tuned rewrite.
"""
import torch
import torch.nn.functional as F


def centers_by_maxpool(heat, thr=0.1, k=7):
    h = F.threshold(heat, thr, -1.)
    pooled = F.max_pool2d(h, kernel_size=k, stride=1, padding=k // 2)
    if k % 2 == 0:
        pooled = pooled[..., :-1, :-1]
    h = h.clone()
    h[h != pooled] = -1.
    h = h.squeeze()
    assert h.dim() == 2
    return torch.nonzero(h > 0)


def nearest_center_chunked(centers, loc, chunk=20):
    n = loc.size(1)
    ids = torch.zeros(n, dtype=torch.long)
    best = 1e5 * torch.ones(n, dtype=torch.float)
    first = 1
    for part in torch.split(centers, chunk, dim=0):
        d = torch.norm(part - loc, dim=-1)
        dmin, amin = d.min(dim=0)
        closer = dmin < best
        ids[closer] = first + amin[closer]
        best = torch.min(best, dmin)
        first += part.size(0)
    return ids


def pixel_ids(centers, offsets, chunk=20, step=1.0):
    assert centers.size(0) > 0
    if offsets.size(0) != 1:
        raise ValueError('Only supports inference for batch size = 1')
    off = offsets.squeeze(0)
    H, W = off.shape[1:]
    ys = torch.arange(0, int(H * step), step=step, dtype=off.dtype).repeat(1, W, 1).transpose(1, 2)
    xs = torch.arange(0, int(W * step), step=step, dtype=off.dtype).repeat(1, H, 1)
    loc = (torch.cat((ys, xs), dim=0) + off).reshape((2, H * W)).transpose(1, 0).unsqueeze(0)
    c = (step * centers).unsqueeze(1)
    if c.size(0) <= chunk:
        ids = 1 + torch.argmin(torch.norm(c - loc, dim=-1), dim=0)
    else:
        ids = nearest_center_chunked(c, loc, chunk)
    return ids.reshape((1, H, W))


def instance_map(sem, heat, offsets, things, thr=0.1, k=7):
    assert sem.size(0) == 1
    sem = sem[0]
    is_thing = torch.zeros_like(sem)
    for t in things:
        is_thing[sem == t] = 1
    ctr = centers_by_maxpool(heat, thr, k)
    if ctr.size(0) == 0:
        return torch.zeros_like(sem), ctr.unsqueeze(0)
    return is_thing * pixel_ids(ctr, offsets), ctr.unsqueeze(0)


def vote_and_paste(sem, ins, divisor, things, stuff_area, void):
    pan = torch.zeros_like(sem) + void
    occupied = ins > 0
    sem_is_thing = torch.zeros_like(sem)
    for t in things:
        sem_is_thing[sem == t] = 1
    next_id = {}
    for i in torch.unique(ins):
        if i == 0:
            continue
        m = (ins == i) & (sem_is_thing == 1)
        if torch.count_nonzero(m) == 0:
            continue
        cls, _ = torch.mode(sem[m].view(-1, ))
        cls_i = cls.item()
        new = next_id.get(cls_i, 1)
        next_id[cls_i] = new + 1
        pan[m] = cls * divisor + new
    for c in torch.unique(sem):
        if c.item() in things:
            continue
        m = (sem == c) & (~occupied)
        if torch.nonzero(m).size(0) >= stuff_area:
            pan[m] = c * divisor
    return pan


def panoptic(sem, heat, offsets, things, divisor, stuff_area, void, thr=0.1, k=7):
    if sem.size(1) != 1:
        raise ValueError('Expect single channel semantic segmentation. Softmax/argmax first!')
    if sem.size(0) != 1 or heat.size(0) != 1 or offsets.size(0) != 1:
        raise ValueError('Only supports inference for batch size = 1')
    ins, ctr = instance_map(sem, heat, offsets, things, thr, k)
    return vote_and_paste(sem, ins, divisor, things, stuff_area, void), ctr
