"""cProfile of StackShard.match on the config-3 stack (one GPU): where the host side of the cross-slice matcher spends
its time.  Usage (GPU box): python profiles/prof_match.py [depth] > gpurun_out/prof_match.txt"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_stack as bs                                                   # noqa: E402


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    slices = bs.make_slices(dev, 2048)
    rec, out, shard, _ = bs.run_stack(dev, 0, 1, slices, depth, 2048, repeats=2, warmup=1)
    for rep in range(2):
        t = time.perf_counter()
        shard.match(out)
        print(f'match (run {rep}): {time.perf_counter() - t:.3f} s for {depth} slices', flush=True)
    pr = cProfile.Profile()
    pr.enable()
    shard.match(out)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
    print(s.getvalue())
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(30)
    print(s.getvalue())


if __name__ == '__main__':
    main()
