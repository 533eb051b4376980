import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lzbench_b200 as L
n = 1024 << 20
h_plain = torch.empty(n, dtype=torch.uint8).pin_memory()
L.corpus("mixed", n, seed=2, out=h_plain.numpy())
ctx = L.Lz4B200(0, n, 65536)
h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
cs = L.chunk_list(n, 65536)
h_out = torch.empty(L.get_compress_bound(n), dtype=torch.uint8).pin_memory()
_, compr, total = ctx.compress_batch(h_plain, cs, h_out)
def dec2():
    ctx.decompress_batch(h_out, compr, cs, h_back)
for K in (16, 32):
    for team in (0, 4, 8, 16, 32):
        ctx.set_option("slices", str(K)); ctx.set_option("team", str(team))
        for _ in range(2): dec2()
        ctx.set_option("timing", "1"); dec2(); kt = ctx.kernel_times_ms(); ctx.set_option("timing", "0")
        t0 = time.perf_counter()
        for _ in range(5): dec2()
        td = (time.perf_counter() - t0) / 5
        print(f"slices={K} team={team}: {td*1e3:.1f} ms; decode kernels: n={len(kt)} sum={sum(kt):.2f} first={kt[0]:.2f} min={min(kt):.2f} max={max(kt):.2f}", flush=True)
