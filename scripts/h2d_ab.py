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
h_frame = torch.empty(L.lib.lz4b200_frame_bound(n, 4), dtype=torch.uint8).pin_memory()
f = ctx.frame_compress(h_plain.numpy(), 4, outbuf=h_frame.numpy())
h_f2 = torch.empty(len(f), dtype=torch.uint8).pin_memory(); h_f2.numpy()[:] = f; f2 = h_f2.numpy()
def t(fn, reps=6):
    for _ in range(2): fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
for rnd in range(3):
    for mode in ("0", "1"):
        ctx.set_option("ramp", mode)
        td = t(lambda: ctx.decompress_batch(h_packed, compr, cs, h_back))
        tc = t(lambda: ctx.compress_batch(h_plain, cs, h_out))
        tfc = t(lambda: ctx.frame_compress(h_plain.numpy(), 4, outbuf=h_frame.numpy()))
        tfd = t(lambda: L.lib.lz4b200_frame_decompress(ctx._h, f2.ctypes.data, f2.nbytes, h_back.numpy().ctypes.data, n, 0))
        print(f"ramp={mode}: decompress {td:.1f} ms  compress {tc:.1f} ms  frame compress {tfc:.1f} ms  frame decompress {tfd:.1f} ms", flush=True)
