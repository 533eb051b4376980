import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lzbench_b200 as L
n = 1024 << 20; chunk = 65536
h_plain = torch.empty(n, dtype=torch.uint8).pin_memory()
L.corpus("mixed", n, seed=2, out=h_plain.numpy())
cs = L.chunk_list(n, chunk)
ctx = L.Lz4B200(0, n, chunk)
h_out = torch.empty(L.get_compress_bound(n), dtype=torch.uint8).pin_memory()
_, compr, total = ctx.compress_batch(h_plain, cs, h_out)
h_packed = torch.empty(total, dtype=torch.uint8).pin_memory(); h_packed.copy_(h_out[:total])
h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
for gen in sys.argv[1].split(","):
    ctx.set_option("kernels", gen)
    for K in [0, 8, 12, 16, 24]:
        ctx.set_option("slices", str(K))
        for _ in range(2): ctx.decompress_batch(h_packed, compr, cs, h_back)
        t0 = time.perf_counter()
        for _ in range(5): ctx.decompress_batch(h_packed, compr, cs, h_back)
        td = (time.perf_counter() - t0) / 5
        for _ in range(2): ctx.compress_batch(h_plain, cs, h_out)
        t0 = time.perf_counter()
        for _ in range(5): ctx.compress_batch(h_plain, cs, h_out)
        tc = (time.perf_counter() - t0) / 5
        print(f"{gen} slices={K}: e2e decompress {n/td/1e9:.1f} GB/s ({td*1e3:.1f} ms)  compress {n/tc/1e9:.1f} GB/s ({tc*1e3:.1f} ms)", flush=True)
assert np.array_equal(h_back.numpy(), h_plain.numpy())
