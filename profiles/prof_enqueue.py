import sys, time, ctypes
sys.path.insert(0, '/root/repo')
import torch
import bench_stack as bs
from empanada_b200.inference import stack
dev = torch.device('cuda', 0)
slices = bs.make_slices(dev, 2048, 12, 400)
# monkeypatch timing into _enqueue_blocks via line profiler-ish wrappers
import empanada_b200._cabi as C
L = C.lib()
orig = L.emp_stack_blocks
tt = {}
def timed(*a):
    t = time.perf_counter(); r = orig(*a); tt['c_call'] = tt.get('c_call', 0) + time.perf_counter() - t; return r
class Wrap:
    def __getattr__(self, k):
        return timed if k == 'emp_stack_blocks' else getattr(L, k)
C._lib = Wrap()
for _ in range(4):
    tt.clear()
    rec, out, shard, _ = bs.run_stack(dev, 0, 1, slices, 64, 2048, repeats=1, warmup=0)
    print(rec['seconds'], shard.timing_, tt)
import cProfile, pstats
eng = bs.make_engine()
def once():
    shard = stack.StackShard(eng, labels=[1], depth=64)
    for i, z in enumerate(shard.slices()):
        s = slices[z % 12]; shard.add(z, s['sem_prob'], s['ctr_hmp'], s['offsets'], size=(2048, 2048))
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable(); out = shard.finish(); pr.disable()
    return pr
once(); pr = once()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
