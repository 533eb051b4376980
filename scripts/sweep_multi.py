"""BASELINE configs[4] on N GPUs of one box (torchrun, one process per GPU): the chunk-size sweep of scripts/sweep.py with the
corpus sharded by contiguous ranges -- rank r owns bytes [r*n, (r+1)*n) -- every rank device-resident on its own shard, no
data-path collective (the path shards trivially; NCCL carries the barrier and the per-rank times).  Rank 0 prints one JSON line
per (corpus, chunk): whole-job GB/s = all ranks' bytes / the slowest rank's time.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/sweep_multi.py [MiB per GPU] [corpora]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import lzbench_b200 as L

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["random", "repeat", "zero", "mixed"]
sizes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]
ctx = L.Lz4B200(local)
s = torch.cuda.current_stream().cuda_stream
for kind in kinds:
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    L.corpus(kind, n, seed=4 if kind == "random" else 5 if kind != "mixed" else 3, offset=rank * n, out=host.numpy())
    d_in = host.cuda()
    d_out = torch.empty(L.get_compress_bound(n) + n // 64, dtype=torch.uint8, device="cuda")
    d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_total = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    for kb in sizes:
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
                if world > 1: dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
                if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the slowest rank
                ts.append(float(t.item()))
            res[name] = min(ts)
        tot = torch.tensor([float(d_total.item()), float(bool(torch.equal(d_back, d_in)) and int(d_bad.item()) == 0)], dtype=torch.float64, device="cuda")
        if world > 1:
            both = [torch.zeros_like(tot) for _ in range(world)]
            dist.all_gather(both, tot)
        else:
            both = [tot]
        total = sum(float(b[0].item()) for b in both); ok = all(bool(b[1].item()) for b in both)
        if rank == 0:
            print(json.dumps({"n_gpus": world, "corpus": kind, "chunk_kib": kb, "chunks_per_gpu": nc, "ratio_pct": round(100 * total / (n * world), 3),
                              "compress_GBps": round(world * n / res["compress"] / 1e6, 1), "decompress_GBps": round(world * n / res["decompress"] / 1e6, 1),
                              "roofline_frac_decompress": round((n * world + total) / res["decompress"] / 1e6 / (6454.6 * world), 4),
                              "roundtrip_ok": ok}), flush=True)
if world > 1: dist.destroy_process_group()
