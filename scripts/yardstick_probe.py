import sys, time; sys.path.insert(0,'/root/repo')
import torch, lzbench_b200 as L
n=1<<30
a=torch.empty(n,dtype=torch.uint8).pin_memory(); b=torch.empty(n,dtype=torch.uint8).pin_memory()
ctx=L.Lz4B200(0,n,65536)
for _ in range(2): ctx.memcpy_batch(a,n,b)
t0=time.perf_counter()
for _ in range(5): ctx.memcpy_batch(a,n,b)
t=(time.perf_counter()-t0)/5
print(f"memcpy_batch: {t*1e3:.1f} ms = {n/t/1e9:.1f} GB/s per direction")
