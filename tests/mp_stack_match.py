"""Multi-GPU check of the z-sharded stack driver incl. cross-rank matching (not collected by pytest):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mp_stack_match.py

Every rank post-processes + RLE-encodes + matches its z-block of the same synthetic stack
(StackShard.finish / match, carry planes and matcher state exchanged over NCCL); rank 0 also runs the
whole stack as a single block and checks that the sharded result — matched labels, boxes, runs — and the
dense fill are identical; the same with the blocks streamed in sub-blocks of 32 slices (StackShard.advance(), on the
current stream and on a side stream), and a table overflow on one rank must raise on all of them."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from empanada_b200.inference import engines, stack  # noqa: E402
from empanada_b200.synth import synth_stack_slices  # noqa: E402


def run(rank, world, heads, D, H, W, ks, dev, group_ok, streamed=False):
    eng = engines.PanopticDeepLabRenderEngine(torch.nn.Identity(), thing_list=[1], label_divisor=20000, stuff_area=64,
                                              void_label=0, nms_threshold=0.1, nms_kernel=3, confidence_thr=0.3)
    shard = stack.StackShard(eng, labels=[1], depth=D, rank=rank, world_size=world, median_kernel_size=ks,
                             block=32 if streamed else 128, stream=torch.cuda.Stream(dev) if streamed == 'side' else None)
    early = 0
    for z in shard.slices():
        if streamed:                                            # sub-blocks of 32 slices leave while the rest is still arriving
            shard.add(z, heads[z]['sem_prob'].clone(), heads[z]['ctr_hmp'].clone(), heads[z]['offsets'].clone(), size=(H, W))
            early += shard.advance()
        else:
            shard.add(z, heads[z]['sem_prob'], heads[z]['ctr_hmp'], heads[z]['offsets'], size=(H, W))
    assert not streamed or early >= 1
    segs = shard.finish()
    matched = shard.match(segs)
    vol = shard.fill(torch.int64).cpu().numpy()
    return matched, vol


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    # 40 slices per rank at 2 ranks, streamed as 32 + 8: two chains over iid noise merge with p ~ 1/2 per pixel and
    # slice, so 24 576 pixels need ~15 slices to settle and 32 leave a 1e-5 chance of the 'did not settle' error
    D, H, W, ks = 80, 128, 192, 3
    sl = list(synth_stack_slices(D, H, W, 40, seed=11, coarse=4, sigma=4.0, z_extent=(5, 16), semi_axes=(6, 22)))
    heads = [{k: torch.from_numpy(s[k]).to(dev) for k in ('sem_prob', 'ctr_hmp', 'offsets')} for s in sl]
    results = []
    for streamed in (False, True, 'side'):
        matched, vol = run(rank, world, heads, D, H, W, ks, dev, True, streamed)
        gathered = [None] * world
        dist.all_gather_object(gathered, (matched, vol))
        results += gathered
    gathered = results
    ok = True
    if rank == 0:
        # single block over the whole stack (world_size=1 code path; no collectives are issued)
        whole, vol1 = run(0, 1, heads, D, H, W, ks, dev, False)
        z0 = 0
        n_obj = set()
        for m, v in gathered:
            for i, z in enumerate(sorted(m)):
                a, b = m[z][1], whole[z][1]
                same = list(a.keys()) == list(b.keys()) and all(
                    tuple(a[k]['box']) == tuple(b[k]['box']) and np.array_equal(a[k]['starts'], b[k]['starts'])
                    and np.array_equal(a[k]['runs'], b[k]['runs']) for k in b)
                same = same and np.array_equal(v[i], vol1[z])
                n_obj |= set(b.keys())
                if not same:
                    ok = False
                    print(f'MISMATCH at slice {z}: {sorted(a)[:6]} vs {sorted(b)[:6]}')
        print(f'mp_stack_match world={world}: {"OK" if ok else "FAILED"} — {D} slices, {len(n_obj)} tracked objects; '
              f'whole-block, streamed and side-stream runs')
    # a table overflow on ANY rank must surface on EVERY rank (the gathered maxima carry it), or the others would
    # wait in the matcher's hand-over: rank 0 alone gets capacities too small for its slices
    eng = engines.PanopticDeepLabRenderEngine(torch.nn.Identity(), thing_list=[1], label_divisor=20000, stuff_area=64,
                                              void_label=0, nms_threshold=0.1, nms_kernel=3, confidence_thr=0.3)
    shard = stack.StackShard(eng, labels=[1], depth=D, rank=rank, world_size=world, median_kernel_size=ks,
                             run_cap=8 if rank == 0 else None, inst_cap=4 if rank == 0 else None)
    for z in shard.slices():
        shard.add(z, heads[z]['sem_prob'], heads[z]['ctr_hmp'], heads[z]['offsets'], size=(H, W))
    raised = False
    try:
        shard.finish()
    except RuntimeError as err:
        raised = 'overflowed' in str(err)
    flags = [None] * world
    dist.all_gather_object(flags, raised)
    if rank == 0:
        print(f'overflow on rank 0 raised on every rank: {all(flags)} {flags}')
        ok = ok and all(flags)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == '__main__':
    main()
