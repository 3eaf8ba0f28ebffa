import torch, statistics
dev=torch.device('cuda',0)
x=torch.empty((16,4096,4096),dtype=torch.int64,device=dev)
y=torch.empty_like(x)
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
gb=x.numel()*8/1e9
for name,fn,bytes_ in [('zero_',lambda: x.zero_(),gb),('fill_',lambda: x.fill_(7),gb),('copy_',lambda: y.copy_(x),2*gb),('sum(read)',lambda: x.sum(),gb)]:
    ms=t(fn); print(f'{name:10s} {ms:.3f} ms  {bytes_/ms*1e3/1e3:.2f} TB/s')
