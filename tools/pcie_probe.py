"""Pinned-memory copy bandwidth of the box (H2D, D2H, both at once): the ceiling of bench.py's `e2e` number.
Measured on the round-1 B200 box: H2D 51 GB/s alone (1 GB transfers), 43 GB/s while a D2H stream runs next to it."""
import torch, time
dev=torch.device('cuda',0)
for mb in (48, 256, 1024):
    n=mb<<20
    h=torch.empty(n,dtype=torch.uint8).pin_memory()
    d=torch.empty(n,dtype=torch.uint8,device=dev)
    for _ in range(2): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize()
    t0=time.perf_counter()
    reps=max(2, 2048//mb)
    for _ in range(reps): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize()
    dt=time.perf_counter()-t0
    print("H2D %4d MB chunks: %.1f GB/s"%(mb, n*reps/dt/1e9))
    h2=torch.empty(n,dtype=torch.uint8).pin_memory()
    t0=time.perf_counter()
    for _ in range(reps): h2.copy_(d,non_blocking=True)
    torch.cuda.synchronize()
    dt=time.perf_counter()-t0
    print("D2H %4d MB chunks: %.1f GB/s"%(mb, n*reps/dt/1e9))
# bidirectional
n=1024<<20
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device=dev)
h2=torch.empty(n//4,dtype=torch.uint8).pin_memory(); d2=torch.empty(n//4,dtype=torch.uint8,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("bidir: H2D %.1f GB/s with D2H %.1f GB/s"%(n*4/dt/1e9, n/4*4/dt/1e9))
