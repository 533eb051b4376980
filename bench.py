#!/usr/bin/env python
"""bench.py -- the LZ4 chunked hot path on B200 (BASELINE.json metric / configs[1]).

A "step" is one pass of the hot path over one batch: decompression of the rank's whole corpus
shard (1 GiB of synthetic "mixed" data in 64 KiB chunks = 16384 independent LZ4 blocks), plus --
reported in the same JSON line under "compress" -- one compression pass of the same shard.

  value      device-resident decompress throughput, GB/s of UNCOMPRESSED bytes (GB = 1e9, as lzbench's
             MB = 1e6, _lzbench/lzbench.cpp:104-105), all ranks' bytes / max-over-ranks time
  e2e        same metric through the reference-facing C-ABI call lz4b200_decompress_batch with pinned
             HOST buffers: H2D of the compressed image + size tables and D2H of the result inside
  roofline   (U + C) bytes per launch of the codec pass / its CUDA-event duration, vs measured HBM peak (decode is
             three kernels since round 2 -- sequence table, order, lane-per-sequence decoder -- timed as ONE event pair:
             the fraction is that of the whole pass, not of its fastest kernel)
  cpu_baseline  the reference's own lz4 1.9.3 (oracle/_ref, compiled from the reference tree) driven by
             lzbench's chunk loop on the host cores, bounded sample

`--impl reference` times that CPU path alone (rank 0 only).  Multi-GPU: one process per GPU, contiguous
chunk ranges per rank (weak scaling: every rank owns a 1 GiB shard of the 8 GiB corpus), one NCCL
all-gather of the shard totals for the global offsets; no data-path collective.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GB = 1e9


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--size-mib", type=int, default=1024, help="uncompressed bytes per GPU")
    p.add_argument("--chunk", type=int, default=65536)
    p.add_argument("--corpus", default="mixed")
    p.add_argument("--cpu-sample-mib", type=int, default=256)
    p.add_argument("--kernels", default=None, help="v0 | v1 (default: library default)")
    return p.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        inside = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.02]
        if inside:
            return inside
        near = sorted(self.rows, key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))   # timed region shorter than
        return [r for (_, r) in near[:3]]                                                 # the sampling period

    def stop(self):
        if self.proc:
            self.proc.terminate()

    @staticmethod
    def summarise(rows):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(plain: np.ndarray, chunk: int, sample_bytes: int, threads: int, steps: int = 1):
    """Reference lz4 (oracle/_ref) -- or the oracle port when _ref is absent -- under lzbench's chunk loop,
    on `threads` host threads, over the first `sample_bytes` of the workload.  bench.py's only use of oracle/."""
    from tests import oracle_lib as O
    import lzbench_b200 as L
    sample = plain[:sample_bytes]
    cs = L.chunk_list(len(sample), chunk)
    stride = L.compress_bound(chunk)
    stride = (stride + 15) & ~15
    packed = np.zeros(len(cs) * stride, dtype=np.uint8)
    compr = np.zeros(len(cs), dtype=np.uint64)
    back = np.zeros(len(sample), dtype=np.uint8)
    tramp = _ref_trampolines()
    if tramp is not None:
        kind = "reference"      # the reference's own lz4.c behind its two plugin wrappers (oracle/ref_plugin.c)
        comp_fn, deco_fn = tramp
    else:
        kind = "port"           # oracle/_ref did not travel: time the C restatement instead
        comp_fn, deco_fn = O.PLUGIN_COMPRESS, O.PLUGIN_DECOMPRESS
    res = {}
    for t in sorted({1, threads}):
        best_c = best_d = None
        for _ in range(steps):
            ns = O.orc.oracle_mt_run(comp_fn, 0, t, cs.ctypes.data, compr.ctypes.data, len(cs),
                                     sample.ctypes.data, packed.ctypes.data, stride)
            assert ns > 0
            best_c = ns if best_c is None else min(best_c, ns)
            ns = O.orc.oracle_mt_run(deco_fn, 1, t, cs.ctypes.data, compr.ctypes.data, len(cs),
                                     back.ctypes.data, packed.ctypes.data, stride)
            assert ns > 0
            best_d = ns if best_d is None else min(best_d, ns)
        assert np.array_equal(back, sample), "CPU baseline round trip failed"
        res[t] = {"compress_GBps": len(sample) / best_c, "decompress_GBps": len(sample) / best_d}
    return {"kind": kind, "threads": res, "compressed_bytes": int(compr.sum()), "sample_bytes": int(len(sample)),
            "compr_sizes": compr}


def lzbench_row(sample: np.ndarray, chunk: int):
    """The stock `lz4 1.9.3` row of the reference's OWN harness (oracle/_ref/lzbench_ref = unmodified lzbench.cpp
    compiled from the reference tree): `lzbench -elz4 -b<chunk KiB>` on the same sample, single thread."""
    import re
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "lzbench_ref")
    if not os.path.exists(exe) or chunk % 1024:
        return None
    d = "/dev/shm" if os.path.isdir("/dev/shm") else None
    try:
        with tempfile.NamedTemporaryFile(suffix=".bin", dir=d, delete=False) as f:
            sample.tofile(f)
            path = f.name
        r = subprocess.run([exe, "-elz4", f"-b{chunk // 1024}", "-t1,1", "-x", path], capture_output=True, text=True, timeout=120)
        for line in (r.stdout + r.stderr).splitlines():
            m = re.match(r"^lz4 1\.9\.3\s+(\d+) MB/s\s+(\d+) MB/s\s+(\d+)\s+([\d.]+)", line)
            if m:
                return {"command": f"lzbench -elz4 -b{chunk // 1024} -t1,1 -x <sample>", "compress_MBps": int(m.group(1)),
                        "decompress_MBps": int(m.group(2)), "compressed_bytes": int(m.group(3)), "ratio_pct": float(m.group(4))}
    except Exception:
        return None
    finally:
        try:
            os.unlink(path)
        except Exception:
            pass
    return None


def _ref_trampolines():
    """C-level lzbench plugin wrappers around oracle/_ref, compiled once into oracle/_ref (git-ignored)."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_plugin.so")
    if not os.path.exists(so):
        return None
    lib = C.CDLL(so)
    return C.cast(lib.ref_plugin_lz4_compress, C.c_void_p), C.cast(lib.ref_plugin_lz4_decompress, C.c_void_p)


def workload_config(args, world):
    """The `config` object of the JSON line: what defines the workload and nothing else, so that both arms print the same one."""
    n = args.size_mib << 20
    return {"workload": f"LZ4 block decompress, {args.chunk} B chunks, {args.size_mib} MiB '{args.corpus}' corpus per GPU "
                        f"(BASELINE configs[1]; configs[3] sharding when n_gpus>1)",
            "chunk_bytes": args.chunk, "chunks_per_gpu": (n + args.chunk - 1) // args.chunk, "corpus": args.corpus,
            "seed": 2 if world == 1 else 3,
            "l2": "inputs larger than L2: a step writes %.0f MB per GPU and reads its compressed image vs 126 MB L2" % (n / 1e6)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import lzbench_b200 as L
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    n = min(args.cpu_sample_mib, args.size_mib) << 20
    plain = L.corpus(args.corpus, n, seed=2 if world == 1 else 3)
    for _ in range(args.warmup):
        cpu_baseline(plain, args.chunk, min(n, 32 << 20), threads)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        r = cpu_baseline(plain, args.chunk, n, threads)
        vals.append(r["threads"][threads])
    wall = time.perf_counter() - t0
    dec = float(np.median([v["decompress_GBps"] for v in vals]))
    com = float(np.median([v["compress_GBps"] for v in vals]))
    line = {
        "impl": "reference", "metric": "lz4_block_decompress_GBps_uncompressed", "value": dec, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (n / GB) / dec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, world),
        "details": {"what": "CPU lz4 1.9.3 (the reference's own lz4.c behind its plugin wrappers) under lzbench's chunk loop, all host threads"},
        "cpu_baseline": {"value": dec, "unit": "GB/s", "cores": threads, "kind": r["kind"],
                         "sample": f"first {n >> 20} MiB of the workload per step, {threads} threads, "
                                   f"contiguous chunk ranges per thread", "compress_GBps": com,
                         "ratio_pct": 100.0 * r["compressed_bytes"] / r["sample_bytes"]},
        "e2e": {"value": dec, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    _emit(line)


def run_b200(args):
    import torch
    import lzbench_b200 as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.size_mib << 20
    chunk = args.chunk
    # corpus: N=1 -> "mixed-1G" seed 2; N>1 -> rank r owns bytes [r*n, (r+1)*n) of "mixed-8G" seed 3
    seed, offset = (2, 0) if world == 1 else (3, rank * n)
    h_plain = torch.empty(n, dtype=torch.uint8).pin_memory()
    L.corpus(args.corpus, n, seed=seed, offset=offset, out=h_plain.numpy())
    cs = L.chunk_list(n, chunk)
    n_chunks = len(cs)
    ctx = L.Lz4B200(local, n, chunk)
    if args.kernels:
        ctx.set_option("kernels", args.kernels)
    stream = torch.cuda.current_stream().cuda_stream

    d_plain = h_plain.to(dev, non_blocking=False)
    d_cs = torch.from_numpy(cs.astype(np.int64)).to(dev)
    d_packed = torch.empty(L.get_compress_bound(n), dtype=torch.uint8, device=dev)
    d_compr = torch.zeros(n_chunks, dtype=torch.int64, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    d_back = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_bad = torch.zeros(1, dtype=torch.int32, device=dev)

    def compress_dev():
        ctx.compress_device(d_plain, d_cs, n_chunks, chunk, d_packed, d_compr, d_total, stream)

    def decompress_dev():
        ctx.decompress_device(d_packed, d_compr, d_cs, n_chunks, chunk, d_back, d_bad, stream)

    # ---- setup (untimed): the compressed image the decompress steps consume ----
    compress_dev()
    torch.cuda.synchronize()
    c_total = int(d_total.item())
    compr = d_compr.cpu().numpy().astype(np.uint64)
    assert int(compr.sum()) == c_total
    h_packed = torch.empty(c_total, dtype=torch.uint8).pin_memory()
    h_packed.copy_(d_packed[:c_total])
    h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(L.get_compress_bound(n), dtype=torch.uint8).pin_memory()

    # multi-GPU exchange step: all-gather of shard totals -> global offsets of each shard (8 B per rank)
    shard_base = 0
    if dist is not None:
        from lzbench_b200 import sharding
        shard_base = sharding.shard_base_offset(sharding.exchange_totals(c_total, device=dev), rank)

    sampler = ClockSampler(local) if rank == 0 else None

    def timed(fn, steps, warmup):
        """device-timed: K launches between barriers, CUDA events on the launching stream, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        ctx.set_option("timing", "1")
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        w1 = time.time()
        ms = e0.elapsed_time(e1)
        kt = ctx.kernel_times_ms()
        ctx.set_option("timing", "0")
        launches = ctx.launches - l0
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, kt, launches, (w0, w1)

    def timed_host(fn, steps, warmup):
        """end-to-end through the host-buffer C-ABI call (synchronous): wall clock, max over ranks."""
        for _ in range(max(1, warmup // 2)):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt / steps

    K, W = args.steps, args.warmup
    # ---- decompress (headline) ----
    d_ms, d_kt, d_launches, d_win = timed(decompress_dev, K, W)
    assert int(d_bad.item()) == 0, "GPU decoder flagged malformed chunks"
    assert torch.equal(d_back, d_plain), "GPU decompression is not bit-exact"
    h_cs = cs
    e2e_d = timed_host(lambda: ctx.decompress_batch(h_packed, compr, h_cs, h_back), max(2, K // 2), W)
    assert np.array_equal(h_back.numpy(), h_plain.numpy()), "host-path decompression is not bit-exact"
    # ---- compress ----
    c_ms, c_kt, c_launches, c_win = timed(compress_dev, K, W)
    e2e_c = timed_host(lambda: ctx.compress_batch(h_plain, h_cs, h_out), max(2, K // 2), W)
    # yardstick: the same U bytes host -> device -> host through the same staging, no kernel (the reference's cudaMemcpy row)
    e2e_m = timed_host(lambda: ctx.memcpy_batch(h_plain, n, h_back), max(2, K // 2), W)

    clocks = None
    if sampler is not None:
        time.sleep(0.15)
        sampler.stop()
        clocks = ClockSampler.summarise(sampler.window(d_win[0], d_win[1]))
        clocks_c = ClockSampler.summarise(sampler.window(c_win[0], c_win[1]))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    U, Cb = float(n), float(c_total)

    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures under profiles/
    # (profiles/README.md), valid for the default workload only (1 GiB mixed, 64 KiB chunks, default kernels)
    default_cfg = args.size_mib == 1024 and chunk == 65536 and args.corpus == "mixed" and not args.kernels
    traffic = {"decompress": 3.5969e9 if default_cfg else None,          # profiles/r02_decompress_v7_ncu_raw.csv: table pass 0.415 + 0.483, lane-per-sequence pass 1.662 + 1.037 GB
               "compress": 1.6950e9 if default_cfg else None}            # profiles/r02_compress_c3_ncu_raw.csv: 1.076 + 0.619 GB

    def roof(kt, which):
        if not kt:
            return None
        t = float(np.mean(kt)) * 1e-3
        ach = (U + Cb) / t / GB
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic[which], "kernel_ms": t * 1e3, "launches_timed": len(kt), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": U + Cb,
                "timed": ("parse_v7_kernel + order_v7_kernel + decompress_v7_kernel + decompress_v5_kernel (work list, empty here), one event pair"
                          if which == "decompress" and not args.kernels and chunk > 16384 and chunk <= 65536 else
                          "compress_c3_kernel" if which == "compress" and not args.kernels and chunk <= 65536 else "the codec kernel of the pass")}

    line = {
        "metric": "lz4_block_decompress_GBps_uncompressed",
        "value": world * U / (d_ms * 1e-3) / GB, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": d_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, world),
        "details": {"compressed_by": "lz4_b200 (standard LZ4 blocks, cross-decode tested)", "compressed_pct": 100.0 * Cb / U,
                    "kernels": args.kernels or "default", "shard_base_offset": shard_base,
                    "bytes_per_step": "reads %.0f MB, writes %.0f MB" % (Cb / 1e6, U / 1e6)},
        "roofline": roof(d_kt, "decompress"),
        "e2e": {"value": world * U / e2e_d / GB, "unit": "GB/s", "h2d_bytes_per_step": int(Cb + 16 * n_chunks),
                "d2h_bytes_per_step": int(U + 4), "ms_per_step": e2e_d * 1e3,
                "api": "lz4b200_decompress_batch (pinned host buffers)",
                "pcie_yardstick": {"value": world * U / e2e_m / GB, "unit": "GB/s", "ms_per_step": e2e_m * 1e3,
                                   "what": "lz4b200_memcpy_batch: the same U bytes host->device->host, no kernel"}},
        "gpu_launches": int(d_launches),
        "clocks": clocks,
        "compress": {"value": world * U / (c_ms * 1e-3) / GB, "unit": "GB/s", "ms_per_step": c_ms,
                     "roofline": roof(c_kt, "compress"), "gpu_launches": int(c_launches), "clocks": clocks_c if clocks else None,
                     "e2e": {"value": world * U / e2e_c / GB, "unit": "GB/s", "h2d_bytes_per_step": int(U + 8 * n_chunks),
                             "d2h_bytes_per_step": int(Cb + 8 * n_chunks + 8), "ms_per_step": e2e_c * 1e3,
                             "api": "lz4b200_compress_batch (pinned host buffers)"}},
    }
    if world == 1:
        threads = os.cpu_count() or 1
        sample = min(n, args.cpu_sample_mib << 20)
        cb = cpu_baseline(h_plain.numpy(), chunk, sample, threads)
        sample_chunks = len(cb["compr_sizes"])
        gpu_sample_bytes = int(compr[:sample_chunks].sum())
        line["cpu_baseline"] = {
            "value": cb["threads"][threads]["decompress_GBps"], "unit": "GB/s", "cores": threads, "kind": cb["kind"],
            "sample": f"first {sample >> 20} MiB of the same corpus, lzbench chunk loop over lz4 1.9.3, "
                      f"contiguous chunk ranges on {threads} threads",
            "single_thread_GBps": cb["threads"][1]["decompress_GBps"],
            "compress_GBps": cb["threads"][threads]["compress_GBps"],
            "compress_single_thread_GBps": cb["threads"][1]["compress_GBps"],
            "ratio_pct": 100.0 * cb["compressed_bytes"] / cb["sample_bytes"]}
        line["cpu_baseline"]["lzbench_row"] = lzbench_row(h_plain.numpy()[:min(sample, 256 << 20)], chunk)
        line["compress"]["ratio_pct"] = 100.0 * Cb / U
        line["compress"]["size_vs_reference"] = gpu_sample_bytes / cb["compressed_bytes"]
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under torchrun), so
    file descriptor 1 points at stderr until the line is ready."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
