import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lzbench_b200 as L
n = 1024 << 20
h_plain = torch.empty(n, dtype=torch.uint8).pin_memory()
L.corpus("mixed", n, seed=2, out=h_plain.numpy())
ctx = L.Lz4B200(0, n, 65536)
h_frame = torch.empty(L.lib.lz4b200_frame_bound(n, 7), dtype=torch.uint8).pin_memory()
h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
f = ctx.frame_compress(h_plain.numpy(), 4, outbuf=h_frame.numpy())
out = h_back.numpy()
cs = L.chunk_list(n, 65536)
h_out = torch.empty(L.get_compress_bound(n), dtype=torch.uint8).pin_memory()
_, compr, total = ctx.compress_batch(h_plain, cs, h_out)
def dec():
    r = L.lib.lz4b200_frame_decompress(ctx._h, f.ctypes.data, f.nbytes, out.ctypes.data, n, 0)
    assert r == n, (r, L.last_error())
def dec2():
    ctx.decompress_batch(h_out, compr, cs, h_back)
for K in (0, 1, 4, 8, 16, 32):
    ctx.set_option("slices", str(K))
    for name, fn in (("frame", dec), ("chunked", dec2)):
        for _ in range(2): fn()
        ctx.set_option("timing", "1")
        t0 = time.perf_counter(); fn(); t1 = time.perf_counter()
        kt = ctx.kernel_times_ms()
        ctx.set_option("timing", "0")
        t0 = time.perf_counter()
        for _ in range(5): fn()
        td = (time.perf_counter() - t0) / 5
        print(f"slices={K} {name}: {td*1e3:.1f} ms; timed call {1e3*(t1-t0):.1f}; decode kernels: n={len(kt)} sum={sum(kt):.2f} ms max={max(kt):.2f}", flush=True)
