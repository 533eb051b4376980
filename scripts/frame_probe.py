"""End-to-end timing of the frame entry points (host buffers, pinned) next to the chunked ones. GPU box only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lzbench_b200 as L
n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1024 << 20
h_plain = torch.empty(n, dtype=torch.uint8).pin_memory()
L.corpus("mixed", n, seed=2, out=h_plain.numpy())
ctx = L.Lz4B200(0, n, 65536)
h_frame = torch.empty(L.lib.lz4b200_frame_bound(n, 7), dtype=torch.uint8).pin_memory()
h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
BIDS = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (4, 5, 6, 7)
FLAGS = ({},) if len(sys.argv) > 3 else ({}, {"block_checksum": True}, {"content_checksum": True})
for bid in BIDS:
    for flags in FLAGS:
        f = ctx.frame_compress(h_plain.numpy(), bid, outbuf=h_frame.numpy(), **flags)
        for _ in range(2): ctx.frame_compress(h_plain.numpy(), bid, outbuf=h_frame.numpy(), **flags)
        t0 = time.perf_counter()
        for _ in range(5): ctx.frame_compress(h_plain.numpy(), bid, outbuf=h_frame.numpy(), **flags)
        tc = (time.perf_counter() - t0) / 5
        f = ctx.frame_compress(h_plain.numpy(), bid, outbuf=h_frame.numpy(), **flags)
        h_f2 = torch.empty(len(f), dtype=torch.uint8).pin_memory(); h_f2.numpy()[:] = f; f = h_f2.numpy()
        out = h_back.numpy()
        def dec():
            r = L.lib.lz4b200_frame_decompress(ctx._h, f.ctypes.data, f.nbytes, out.ctypes.data, n, 0)
            assert r == n, (r, L.last_error())
        for _ in range(2): dec()
        t0 = time.perf_counter()
        for _ in range(5): dec()
        td = (time.perf_counter() - t0) / 5
        assert np.array_equal(out, h_plain.numpy())
        print(f"block_id={bid} {flags}: frame {len(f)} B ratio {len(f)/n:.4f}  compress {n/tc/1e9:.1f} GB/s ({tc*1e3:.1f} ms)  "
              f"decompress {n/td/1e9:.1f} GB/s ({td*1e3:.1f} ms)", flush=True)
