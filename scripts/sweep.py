"""BASELINE configs[4]: chunk-size sweep 4 KiB - 4 MiB on incompressible and highly repetitive inputs (+ mixed),
device-resident, one GPU.  Prints one JSON line per (corpus, chunk)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lzbench_b200 as L

n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["random", "repeat", "zero", "mixed"]
ctx = L.Lz4B200(0)
s = torch.cuda.current_stream().cuda_stream
for kind in kinds:
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    L.corpus(kind, n, seed=4 if kind == "random" else 5 if kind != "mixed" else 2, out=host.numpy())
    d_in = host.cuda()
    d_out = torch.empty(L.get_compress_bound(n) + n // 64, dtype=torch.uint8, device="cuda")
    d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_total = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    for kb in [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]:
        chunk = kb << 10
        cs = L.chunk_list(n, chunk).astype(np.int64)
        nc = len(cs)
        d_cs = torch.from_numpy(cs).cuda()
        d_compr = torch.empty(nc, dtype=torch.int64, device="cuda")
        res = {}
        for name, fn in (("compress", lambda: ctx.compress_device(d_in, d_cs, nc, chunk, d_out, d_compr, d_total, s)),
                         ("decompress", lambda: ctx.decompress_device(d_out, d_compr, d_cs, nc, chunk, d_back, d_bad, s))):
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[name] = min(ts)
        total = int(d_total.item())
        ok = bool(torch.equal(d_back, d_in)) and int(d_bad.item()) == 0
        print(json.dumps({"corpus": kind, "chunk_kib": kb, "chunks": nc, "ratio_pct": round(100 * total / n, 3),
                          "compress_GBps": round(n / res["compress"] / 1e6, 1), "decompress_GBps": round(n / res["decompress"] / 1e6, 1),
                          "roofline_frac_decompress": round((n + total) / res["decompress"] / 1e6 / 6454.6, 4),
                          "roundtrip_ok": ok}), flush=True)
