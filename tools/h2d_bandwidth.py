"""Bare pinned host -> device copy bandwidth of this box (tool): the ceiling of bench.py's end-to-end leg."""
import torch, time
dev=torch.device('cuda:0')
for mb in (92, 370, 1480):
    n=mb*1024*1024//4
    h=torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
    d=torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(mb, 'MB: %.1f GB/s' % (10*n*4/ (e0.elapsed_time(e1)*1e-3)/1e9))
