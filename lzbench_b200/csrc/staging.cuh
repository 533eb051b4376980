// staging.cuh -- chunk staging kernels: exclusive scan of chunk sizes and the compaction
// pass that turns bound-strided compressed slots into the contiguous image lzbench
// expects (`outbuf += clen`, _lzbench/lzbench.cpp:291-294).  The reference's own GPU codec
// does the same in three kernels (nvcomp/LZ4CompressionKernels.cu:1038-1097).
#pragma once
#include "common.cuh"

namespace lz4b200 {

constexpr int SCAN_THREADS = 1024;

// Exclusive prefix sum of up to two u64 arrays in one launch (blockIdx.x picks the array).
// One CTA per array: every thread owns a contiguous run, the run totals are scanned in
// shared memory, then the run is rewritten.  n is at most a few million (8 GiB / 4 KiB).
struct ScanJob { const u64* sizes; u64* offsets; u64* total; const u64* base; u64* mirror; };   // offsets start at *base (0 if null); *total = *base + sum, also stored to *mirror (host-mapped memory: the host reads it after an event, no copy engine involved)

__global__ void __launch_bounds__(SCAN_THREADS) scan_sizes_kernel(ScanJob j0, ScanJob j1, u64 n)
{
    const ScanJob j = blockIdx.x == 0 ? j0 : j1;
    if (j.sizes == nullptr) return;
    __shared__ u64 warp_sums[32];
    __shared__ u64 carry;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const u64 lo = (u64)tid * per < n ? (u64)tid * per : n;
    const u64 hi = lo + per < n ? lo + per : n;
    u64 local = 0;
    for (u64 i = lo; i < hi; i++) local += j.sizes[i];
    // inclusive scan of `local` across the block
    u64 x = local;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 y = __shfl_up_sync(FULL, x, d);
        if (lane >= (u32)d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        u64 w = warp_sums[lane];
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 y = __shfl_up_sync(FULL, w, d);
            if (lane >= (u32)d) w += y;
        }
        warp_sums[lane] = w;
        if (lane == 31) carry = w;
    }
    __syncthreads();
    const u64 start = j.base ? *j.base : 0;
    u64 base = start + x - local + (warp ? warp_sums[warp - 1] : 0);
    for (u64 i = lo; i < hi; i++) { u64 s = j.sizes[i]; j.offsets[i] = base; base += s; }
    if (tid == 0 && j.total) *j.total = start + carry;
    if (tid == 0 && j.mirror) *j.mirror = start + carry;
}

// Compaction: one CTA per chunk (grid-stride).  Source is the chunk's staging slot, or the
// uncompressed input when the block came out exactly chunk-sized (lzbench stores such a
// chunk raw, _lzbench/lzbench.cpp:284-289, and the decoder then copies it back raw, :311-315).
constexpr int COMPACT_THREADS = 256;

__global__ void __launch_bounds__(COMPACT_THREADS)
compact_kernel(const u8* __restrict__ stage, u64 slot_stride, const u8* __restrict__ in,
               const u64* __restrict__ in_off, const u64* __restrict__ usz,
               const u64* __restrict__ csz, const u64* __restrict__ out_off,
               u8* __restrict__ out, u32 n_chunks)
{
    for (u32 c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        u64 len = csz[c];
        const u8* src = (len == usz[c]) ? in + in_off[c] : stage + (u64)c * slot_stride;
        coop_copy(out + out_off[c], src, len, threadIdx.x, COMPACT_THREADS);
    }
}

}  // namespace lz4b200
