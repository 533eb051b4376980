"""Quick device-resident timing probe (not the bench): GB/s of uncompressed bytes per kernel generation."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import lzbench_b200 as L

n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1024 << 20
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
kinds = sys.argv[3].split(",") if len(sys.argv) > 3 else ["mixed"]
gens = sys.argv[4].split(",") if len(sys.argv) > 4 else ["v0", "v1"]

ctx = L.Lz4B200(0)
for kind in kinds:
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    t0 = time.time()
    if kind.startswith("files:"):          # real data: concatenated files (shared objects of the image), repeated to n bytes
        import glob
        buf = bytearray()
        for f in sorted(glob.glob(kind[6:])):
            if os.path.isfile(f) and not os.path.islink(f):
                buf += open(f, "rb").read()
            if len(buf) >= min(n, 128 << 20): break
        a = np.frombuffer(bytes(buf), dtype=np.uint8)
        reps = -(-n // len(a))
        host.numpy()[:] = np.tile(a, reps)[:n] if reps > 1 else a[:n]
        print("   real-file corpus:", len(a) >> 20, "MiB of distinct data", flush=True)
    else:
        L.corpus(kind, n, seed=2, out=host.numpy())
    tg = time.time() - t0
    d_in = host.cuda()
    n_chunks = (n + chunk - 1) // chunk
    cs = L.chunk_list(n, chunk).astype(np.int64)
    d_cs = torch.from_numpy(cs).cuda()
    d_out = torch.empty(L.get_compress_bound(n), dtype=torch.uint8, device="cuda")
    d_compr = torch.empty(n_chunks, dtype=torch.int64, device="cuda")
    d_total = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for gen in gens:
        ctx.set_option("accel", "1"); ctx.set_option("split", "-1"); ctx.set_option("planes", "32")
        if gen.startswith("v7s"):                        # v7s<percent of a large batch on the team decoder>
            ctx.set_option("kernels", "v7"); ctx.set_option("split", gen[3:])
        elif gen.startswith("v7p"):                      # v7p<chunks per warp in the table pass>
            ctx.set_option("kernels", "v7"); ctx.set_option("planes", gen[3:])
        elif gen.startswith("c2a"):                       # c2a<acceleration>
            ctx.set_option("kernels", "v5"); ctx.set_option("compress", "v2"); ctx.set_option("accel", gen[3:])
        elif gen.startswith("c3"):                       # c3[w<worker warps>]
            ctx.set_option("kernels", "default"); ctx.set_option("compress", "v3")
        elif gen in ("c1", "c2"):
            ctx.set_option("kernels", "v4"); ctx.set_option("compress", "v" + gen[1])
        elif gen.startswith("c1m"):
            ctx.set_option("kernels", "v2"); ctx.set_option("c1mode", gen[3:])
        elif gen.startswith("v6d"):                      # v6d<depth>t<team>
            ctx.set_option("kernels", "v6"); ctx.set_option("depth", gen[3]); ctx.set_option("team", gen[5:])
        elif gen.startswith("v4t") or gen.startswith("v5t") or gen.startswith("v6t"):
            ctx.set_option("kernels", gen[:2]); ctx.set_option("team", gen[3:])
        elif gen.startswith("v3t"):
            ctx.set_option("kernels", "v3"); ctx.set_option("team", gen[3:])
        elif gen.startswith("v2t"):
            ctx.set_option("kernels", "v2"); ctx.set_option("team", gen[3:])
        else:
            ctx.set_option("kernels", gen)
        def comp(): ctx.compress_device(d_in, d_cs, n_chunks, chunk, d_out, d_compr, d_total, s)
        def deco(): ctx.decompress_device(d_out, d_compr, d_cs, n_chunks, chunk, d_back, d_bad, s)
        res = {}
        for name, fn in (("compress", comp), ("decompress", deco)):
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[name] = min(ts)
        if gen == "v1" and os.environ.get("PROF"):
            ctx.set_option("prof", "1"); deco(); pr = ctx.debug_prof(); ctx.set_option("prof", "0")
            tot = sum(pr[:8]) or 1
            print("   decode phases [ticket,load,parse,scan,literals,matches,(coop),store] %:",
                  [round(100 * x / tot, 1) for x in pr[:7]], "cycles/chunk", tot // n_chunks, flush=True)
        if gen.startswith("c3") and os.environ.get("PROF"):
            ctx.set_option("prof", "1"); comp(); pr = ctx.debug_prof(); ctx.set_option("prof", "0")
            tiles = pr[4] or 1
            print("   c3 cycles/tile: H+V %d | T %d | P %d | E %d | step %d | stage/chunk %d | tiles/chunk %.1f"
                  % (pr[0] // tiles, pr[1] // tiles, pr[2] // tiles, pr[3] // tiles, pr[6] // tiles, pr[5] // n_chunks, tiles / n_chunks), flush=True)
            print("   c3 parse: second warp %d cycles/tile; rounds/tile %.2f, steps/tile %.1f (both warps)" % (pr[9] // tiles, pr[7] / tiles, pr[8] / tiles), flush=True)
        elif gen in ("c2", "v5t8") and os.environ.get("PROF"):
            ctx.set_option("prof", "1"); comp(); pr = ctx.debug_prof(); ctx.set_option("prof", "0")
            tiles = n_chunks * 16
            print("   c2 cycles/tile: A match %d, A wait-empty %d | B wait-full %d, B parse %d, B emit %d | stage/chunk %d | rounds/tile %.2f"
                  % (pr[0] // tiles, pr[1] // tiles, pr[2] // tiles, pr[3] // tiles, pr[4] // tiles, pr[5] // n_chunks, pr[7] / tiles), flush=True)
        elif gen != "v0" and os.environ.get("PROF"):
            ctx.set_option("prof", "1"); comp(); pr = ctx.debug_prof(); ctx.set_option("prof", "0")
            tot = sum(pr[8:13]) or 1
            print("   compress phases [load,match,parse,emit,final] %:", [round(100 * x / tot, 1) for x in pr[8:13]],
                  "cycles/chunk", tot // n_chunks, "parse rounds/tile %.2f" % (pr[13] / (n_chunks * 16)), flush=True)
        total = int(d_total.item())
        if os.environ.get("REF"):
            from tests import oracle_lib as O
            m = min(n_chunks, (16 << 20) // chunk)
            acc = int(gen[3:]) if gen.startswith("c2a") else 1
            hb = host.numpy(); tot_ref = 0; tmp = np.empty(L.compress_bound(chunk), dtype=np.uint8)
            for i in range(m):
                part = hb[i * chunk:(i + 1) * chunk]
                r_ = O.orc.oracle_lz4_compress_fast(part.ctypes.data, len(part), tmp.ctypes.data, len(tmp), acc)
                tot_ref += min(r_, len(part))
            print("   size vs reference LZ4_compress_fast(accel=%d) on the first %d chunks: %.4f (reference ratio %.2f%%)"
                  % (acc, m, float(d_compr[:m].sum().item()) / tot_ref, 100.0 * tot_ref / (m * chunk)), flush=True)
        ok = bool(torch.equal(d_back, d_in)) and int(d_bad.item()) == 0
        print(f"{kind} {n>>20}MiB chunk={chunk} {gen}: ratio={100*total/n:.2f}% "
              f"compress {n/res['compress']/1e6:.1f} GB/s ({res['compress']:.2f} ms) "
              f"decompress {n/res['decompress']/1e6:.1f} GB/s ({res['decompress']:.2f} ms) "
              f"roundtrip_ok={ok} (corpus gen {tg:.1f}s)", flush=True)
