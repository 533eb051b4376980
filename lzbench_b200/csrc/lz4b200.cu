// lz4b200.cu -- context, chunk staging and the C ABI (include/lz4b200.h) of the B200-native
// LZ4 block codec.  Host logic mirrors the reference's two chunk loops
// (_lzbench/lzbench.cpp:266-298 lzbench_compress, :301-329 lzbench_decompress) as ONE
// batched pass over all chunks; the reference's own GPU plugin
// (_lzbench/compressors.cpp:1856-2010) is the precedent for owning device buffers behind
// `workmem`.  There is no CPU codec in this file: if CUDA is unavailable every entry point
// fails loudly.
#include "../../include/lz4b200.h"
#include "common.cuh"
#include "kernels_v0.cuh"
#include "kernels_v1.cuh"
#include "kernels_v2.cuh"
#include "kernels_v3.cuh"
#include "kernels_v4.cuh"
#include "kernels_v5.cuh"
#include "kernels_v6.cuh"
#include "kernels_v7.cuh"
#include "kernels_c1.cuh"
#include "kernels_c2.cuh"
#include "kernels_c3.cuh"
#include "kernels_coarse.cuh"
#include "segmented.cuh"
#include "frame.cuh"
#include "staging.cuh"

#include <algorithm>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

using namespace lz4b200;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    if (getenv("LZ4B200_VERBOSE")) fprintf(stderr, "lz4b200: %s\n", g_err);
    return code;
}

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(-2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = n + (n >> 3) + 4096;                  // slack: avoid regrowing on jitter
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = n + (n >> 3) + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct lz4b200_ctx {
    int device = 0;
    int sm_count = 148;
    int kernel_gen = 7;                 // decode: 7 = sequence table + one lane per sequence (kernels_v7.cuh; default for chunks <= 64 KiB, v5 otherwise and for the chunks v7 leaves to it); 0 warp per chunk, 1 CTA tile (TMA), 2 sub-warp teams, 3 + mapped lanes / pipelined match, 4 + warp-uniform fast path, 5 + compressed stream through a cp.async ring, long literal runs on the common path (default), 6 = 5 with cp.async-staged deep match pipeline (experiment)
    int compress_gen = 3;               // compress: 0 warp per chunk, 1 CTA per chunk (parallel match finder + speculative parse), 2 = 1 with the phases on two warp groups,
                                        // 3 (default) = kernels_c3.cuh for chunks up to 64 KiB at acceleration 1 (five stages on their own warps), 2 for everything else
    int c1mode = 2;                     // compress: intra-warp candidate mode of the match finder (experiments; see kernels_c1.cuh)
    int v6depth = 2;                    // decode v6: iterations between fetching and storing a match part
    int accel = 1;                      // compress: LZ4_compress_fast's acceleration (1 = lz4 default; lz4fast levels), probe grid 1/2/4/8
    int team = 0;                       // lanes per chunk for the team decoders: 8 / 16 / 32, 0 = by batch size
    int slices = 0;                     // host entry points: number of pipeline slices (0 = automatic)
    bool ramp = true;                   // decompress slice plan: small first slices (option "ramp")
    bool own_h2d = true;                // inbound bulk copies on their own stream (option "h2d" = "own" / "shared")
    bool host_register = false;         // page-lock caller buffers on first sight (lzbench's are plain malloc)
    struct Reg { const char* p; size_t n; };
    Reg regs[16] = {};
    int n_regs = 0;
    size_t seg_chunks = 0;              // chunk count the segmented job arrays are laid out for (current call)
    static constexpr int NS = 8;        // compute/H2D streams of the host pipeline
    static constexpr int MAX_SLICES = 64;
    cudaStream_t streams[NS] = {};
    cudaStream_t d2h_stream = nullptr;
    cudaStream_t aux_stream = nullptr;  // v7 decode: the team decoder's share of a large batch runs beside the sequence-table pass
    cudaEvent_t ev_fork[MAX_SLICES] = {}, ev_join[MAX_SLICES] = {};
    size_t c3_min_chunk = 0;            // c3 for every chunk size up to 64 KiB: below 16 KiB it is 2-4 % slower than c2 (a chunk costs 4 pipeline steps beside its
                                        // tiles) but stays within a fraction of a percent of the reference's size on ANY input, where c2's 256-position batches
                                        // lose up to 9 % on small chunks that mix runs and text (scripts/fuzz_gpu.py, same seed: 32 size failures with c2, 17 with c3 --
                                        // the rest are chunks above 64 KiB, whose segments see 16 KiB of history)
    bool c3_forced = false;
    int v7_lanes = 32;                  // v7 pass 1: chunks per warp (option "planes")
    int split = -1;                     // percent of a large batch given to the team decoder (-1 = automatic, option "split")
    cudaStream_t h2d_stream = nullptr;  // ALL inbound bulk copies: spread over the compute streams they lose a third of the link (measured: 1 GiB each way in 32 ms on 8 streams, 22.6 ms on one)
    cudaEvent_t ev_sizes = nullptr, ev_scan[MAX_SLICES] = {}, ev_done[MAX_SLICES] = {}, ev_in[MAX_SLICES] = {};
    uint64_t launches = 0;
    // optional per-launch timing of the dominant (codec) kernel: event pairs on the launching stream
    bool timing = false;
    static constexpr int MAX_TIMED = 256;
    cudaEvent_t ev0[MAX_TIMED] = {}, ev1[MAX_TIMED] = {};
    int n_timed = 0;
    // device-resident scratch shared by both directions
    DevBuf usz, csz, uoff, coff;        // u64[n_chunks]
    DevBuf fsz, fhdr;                   // frame container: u64 framed block sizes / u32 block header words
    DevBuf scalars;                     // see Scalars below
    DevBuf status;                      // u32[n_chunks]
    DevBuf prof;                        // u64[16] phase-cycle counters (debug option "prof")
    bool prof_on = false;
    DevBuf stage;                       // compress: n_chunks slots of slot_stride bytes (or one slot per 64 KiB segment)
    DevBuf seg64, seg32;                // segmented compress (segmented.cuh): u64 / u32 job arrays
    DevBuf coarse;                      // coarse decode (kernels_coarse.cuh): list counters per slice, descriptors, descriptor counts, work list
    size_t coarse_chunks = 0;           // chunk count its arrays are laid out for (current call; 0: not available)
    bool coarse_on = true;              // option "coarse"
    bool v7_big = true;                 // option "v7big": chunks above 64 KiB (up to 16 MiB) through the sequence table as well
    cudaEvent_t ev_coarse = nullptr;    // the last coarse copy kernel of this context (they run one at a time)
    DevBuf v7_table, v7_u32;            // v7 decode: sequence table (u64 entries) and nseq / route / team list (u32[3][n_chunks])
    size_t v7_chunks = 0;               // chunk count the u32 arrays are laid out for (current call)
    size_t v7_table_chunks = 0;         // chunks the table has room for (current call): all of a host call's, one sub-batch of a device call's
    size_t v7_per_chunk = 0;            // table entries reserved per chunk (current call)
    size_t v7_min_chunk = 16385;        // v7 pays on chunks above 16 KiB (1 GiB mixed, v7 vs v5: 8 KiB 275 vs 317 GB/s, 16 KiB 302 vs 299, 32 KiB 327 vs 261, 64 KiB 316 vs 230); "kernels=v7" forces it
    // buffers behind the host entry points
    DevBuf d_plain, d_packed;
    PinBuf h_sizes;                     // u64[2*n_chunks + MAX_SLICES + 2]
    u64* h_mirror = nullptr;            // u64[MAX_SLICES] host-mapped: running totals the scan kernels store for the host
    u64* d_mirror = nullptr;            // the same memory as the device addresses it
};

namespace {

// layout of ctx->scalars (8 KB)
constexpr size_t SC_TOTAL = 0;          // u64  total compressed size (device entry point without caller buffer)
constexpr size_t SC_BAD = 8;            // u32  malformed-chunk counter
constexpr size_t SC_FBASE = 16;         // u64  frame container: position of the first block (= header bytes)
constexpr size_t SC_FUNSUP = 24;        // u32  frame container: blocks the size rule cannot represent
constexpr size_t SC_FCK = 28;           // u32  frame container: block checksum mismatches
constexpr size_t SC_TICKETS = 256;      // u32[MAX_SLICES] work tickets, one per slice
constexpr size_t SC_TICKETS2 = 512;     // u32[MAX_SLICES] second ticket per slice (two decoders side by side)
constexpr size_t SC_NJOBS = 1024;        // u64[MAX_SLICES] job counts of the segmented compressor, one per slice
constexpr size_t SC_TEAMCOUNT = 2048;    // (unused since the v7 zero block)
constexpr size_t SC_RUNNING = 4096;     // u64[MAX_SLICES] running compressed totals after each slice
constexpr size_t SC_BYTES = 8192;

// per-slice view of the scratch arrays: slices of one host call run concurrently on different streams
struct Scratch {
    u64* uoff; u64* coff; u32* status; u8* stage; u32* ticket;
    u64* fsz;                          // frame mode only
    u64* v7_table; u64 v7_cap; u32 *v7_nseq, *v7_route, *v7_list, *v7_count, *v7_order, *v7_bucket;   // kernels_v7.cuh (null: not available)
    u32* ticket2; int slice;
    size_t first_chunk;                // slices of one call own disjoint regions of the per-chunk / per-job arrays
};

void time_begin(lz4b200_ctx* c, cudaStream_t st)
{
    if (!c->timing || c->n_timed >= lz4b200_ctx::MAX_TIMED) return;
    int i = c->n_timed;
    if (!c->ev0[i]) { cudaEventCreate(&c->ev0[i]); cudaEventCreate(&c->ev1[i]); }
    cudaEventRecord(c->ev0[i], st);
}

void time_end(lz4b200_ctx* c, cudaStream_t st)
{
    if (!c->timing || c->n_timed >= lz4b200_ctx::MAX_TIMED) return;
    cudaEventRecord(c->ev1[c->n_timed++], st);
}

size_t slot_stride_for(size_t max_chunk)
{
    size_t b = max_chunk + max_chunk / 255 + 16;
    return (b + 15) & ~size_t(15);
}

bool segmented(const lz4b200_ctx* c, size_t max_chunk) { return c->compress_gen >= 1 && max_chunk > SEG_BYTES; }
size_t segs_per_chunk(size_t max_chunk) { return (max_chunk + SEG_NEW - 1) / SEG_NEW; }

// bytes of staging the compressor needs for n chunks of at most max_chunk bytes
size_t stage_bytes_for(const lz4b200_ctx* c, size_t n, size_t max_chunk)
{
    if (segmented(c, max_chunk)) return n * segs_per_chunk(max_chunk) * slot_stride_for(SEG_BYTES);
    return n * slot_stride_for(max_chunk);
}

int ensure_segmented(lz4b200_ctx* c, size_t n, size_t max_chunk)
{
    const size_t jobs = n * segs_per_chunk(max_chunk);
    CU(c->seg64.ensure((2 * n + 3 * jobs + 8) * sizeof(u64)));
    CU(c->seg32.ensure((6 * jobs + 8) * sizeof(u32)));
    return 0;
}

int ensure_scratch(lz4b200_ctx* c, size_t n, size_t stage_bytes)
{
    CU(c->uoff.ensure(n * sizeof(u64)));
    CU(c->coff.ensure(n * sizeof(u64)));
    CU(c->status.ensure(n * sizeof(u32)));
    if (stage_bytes) CU(c->stage.ensure(stage_bytes));
    if (!c->scalars.p) { CU(c->scalars.ensure(SC_BYTES)); CU(cudaMemset(c->scalars.p, 0, SC_BYTES)); }
    return 0;
}

// v7 decode scratch: u32 arrays for n chunks, table room for table_chunks chunks of at most max_chunk bytes.  Failing to
// get the memory is not an error: the slice then runs on the team decoder.
constexpr size_t V7_SUB = 65536;            // chunks per sub-batch of a device-resident call (bounds the table: 2 x their bytes)
void ensure_v7(lz4b200_ctx* c, size_t n, size_t table_chunks, size_t max_chunk)
{
    c->v7_chunks = c->v7_table_chunks = 0;
    if (c->kernel_gen != 7 || max_chunk > V7_BIG_MAX || (max_chunk > V7_MAX_CHUNK && !c->v7_big) || max_chunk < c->v7_min_chunk || n == 0) return;
    c->v7_per_chunk = (size_t)v7_entries_for(max_chunk, 1) + 2;      // + 2: the floors of a ragged prefix sum
    if (c->v7_u32.ensure((4 * n + V7_ZERO_WORDS * lz4b200_ctx::MAX_SLICES) * sizeof(u32)) != cudaSuccess ||
        c->v7_table.ensure(table_chunks * c->v7_per_chunk * sizeof(u64)) != cudaSuccess) { cudaGetLastError(); return; }
    c->v7_chunks = n; c->v7_table_chunks = table_chunks;
}

// coarse decode scratch for n chunks (chunks above 64 KiB only).  Failing to get the memory is not an error.
void ensure_coarse(lz4b200_ctx* c, size_t n, size_t max_chunk)
{
    c->coarse_chunks = 0;
    if (!c->coarse_on || c->kernel_gen < 5 || max_chunk <= V7_MAX_CHUNK || n == 0) return;
    if (c->coarse.ensure(1024 + n * (C_MAXD * (sizeof(uint4) + sizeof(uint2)) + 2 * sizeof(u32))) != cudaSuccess) { cudaGetLastError(); return; }
    c->coarse_chunks = n;
}

Scratch scratch_at(lz4b200_ctx* c, size_t first_chunk, size_t stride, int slice)
{
    Scratch s;
    s.uoff = (u64*)c->uoff.p + first_chunk;
    s.coff = (u64*)c->coff.p + first_chunk;
    s.status = (u32*)c->status.p + first_chunk;
    s.stage = c->stage.p ? (u8*)c->stage.p + first_chunk * stride : nullptr;   // `stride` = staging bytes per chunk
    s.ticket = (u32*)((u8*)c->scalars.p + SC_TICKETS) + slice;
    s.ticket2 = (u32*)((u8*)c->scalars.p + SC_TICKETS2) + slice;
    s.slice = slice;
    s.fsz = c->fsz.p ? (u64*)c->fsz.p + first_chunk : nullptr;
    s.v7_table = nullptr; s.v7_cap = 0; s.v7_nseq = s.v7_route = s.v7_list = s.v7_count = s.v7_order = s.v7_bucket = nullptr;
    if (c->v7_table.p && c->v7_chunks && first_chunk < c->v7_chunks) {
        // every slice owns the table region of its chunks (slices of one host call run concurrently)
        if (first_chunk < c->v7_table_chunks) {
            s.v7_table = (u64*)c->v7_table.p + first_chunk * c->v7_per_chunk;
            s.v7_cap = (c->v7_table_chunks - first_chunk) * c->v7_per_chunk;
        }
        u32* b = (u32*)c->v7_u32.p;
        s.v7_nseq = b + first_chunk; s.v7_route = b + c->v7_chunks + first_chunk; s.v7_list = b + 2 * c->v7_chunks + first_chunk;
        s.v7_order = b + 3 * c->v7_chunks + first_chunk; s.v7_bucket = b + 4 * c->v7_chunks + V7_ZERO_WORDS * slice;   // the slice's zero block
        s.v7_count = s.v7_bucket + 2 * V7_BUCKETS;
    }
    s.first_chunk = first_chunk;
    return s;
}

// ---- device pipelines (one slice) ---------------------------------------------------------
// Compress chunks [0, n) of d_in (sizes d_usz) into staging slots, scan the sizes starting at *base
// (0 if null), compact into d_out at those absolute offsets, write the running total to *d_total.
int compress_slice(lz4b200_ctx* c, const u8* d_in, const u64* d_usz, size_t n, size_t max_chunk, u8* d_out,
                   u64* d_csz, const u64* base, u64* d_total, const Scratch& s, cudaStream_t st,
                   cudaEvent_t wait_before_scan, cudaEvent_t record_after_scan, int frame = 0, u64* mirror = nullptr)
{
    // frame: 0 = lzbench's contiguous image, 1 = LZ4 frame blocks (LE32 size word before each), 2 = 1 + block checksums
    const size_t stride = slot_stride_for(max_chunk);
    ScanJob in_scan{ d_usz, s.uoff, nullptr, nullptr }, none{ nullptr, nullptr, nullptr, nullptr };
    scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>(in_scan, none, (u64)n);
    c->launches++;

    SegArrays sa{};
    const bool seg = segmented(c, max_chunk);
    time_begin(c, st);
    if (seg) {
        // chunks > 64 KiB: one job per 64 KiB segment on the tiled compressor, stitched afterwards (segmented.cuh)
        const size_t per = segs_per_chunk(max_chunk);
        const size_t jobs_max = n * per, job0 = s.first_chunk * per;
        const size_t n_tot = c->seg_chunks, jobs_tot = n_tot * per;
        u64* p64 = (u64*)c->seg64.p; u32* p32 = (u32*)c->seg32.p;
        sa.jobs_per_chunk = p64 + s.first_chunk;
        sa.job_base = p64 + n_tot + s.first_chunk;
        sa.job_in_off = p64 + 2 * n_tot + job0;
        sa.job_len = p64 + 2 * n_tot + jobs_tot + job0;
        sa.job_csz = p64 + 2 * n_tot + 2 * jobs_tot + job0;
        sa.n_jobs = (u64*)((u8*)c->scalars.p + SC_NJOBS) + (s.ticket - (u32*)((u8*)c->scalars.p + SC_TICKETS));
        sa.job_tail_pos = p32 + job0;
        sa.job_tail_len = p32 + jobs_tot + job0;
        sa.job_rel = p32 + 2 * jobs_tot + job0;
        sa.job_carry = p32 + 3 * jobs_tot + job0;
        sa.job_chunk = p32 + 4 * jobs_tot + job0;
        sa.job_hist = p32 + 5 * jobs_tot + job0;
        const unsigned gb = (unsigned)((n + 255) / 256);
        seg_count_kernel<<<gb, 256, 0, st>>>(d_usz, (u32)n, sa.jobs_per_chunk);
        ScanJob js{ sa.jobs_per_chunk, sa.job_base, sa.n_jobs, nullptr };
        scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>(js, none, (u64)n);
        seg_fill_kernel<<<gb, 256, 0, st>>>(d_usz, s.uoff, sa, (u32)n);
        if (c->compress_gen == 3 && c->accel == 1)
            CU(c3_launch_compress(d_in, sa.job_in_off, sa.job_len, s.stage, slot_stride_for(SEG_BYTES), sa.job_csz,
                                  (u32)jobs_max, s.ticket, c->sm_count, st, nullptr, sa.n_jobs, sa.job_tail_pos, sa.job_tail_len, sa.job_hist));
        else
        CU(c2_launch_compress(d_in, sa.job_in_off, sa.job_len, s.stage, slot_stride_for(SEG_BYTES), sa.job_csz,
                              (u32)jobs_max, s.ticket, c->sm_count, st, sa.n_jobs, sa.job_tail_pos, sa.job_tail_len, sa.job_hist, nullptr,
                              (u32)c->accel, c->c1mode));
        seg_stitch_size_kernel<<<gb, 256, 0, st>>>(sa, s.stage, slot_stride_for(SEG_BYTES), d_csz, (u32)n);
        c->launches += 5;
    } else if (c->compress_gen == 3 && c->accel == 1 && c1_compress_supported(max_chunk) && (max_chunk >= c->c3_min_chunk || c->c3_forced)) {
        // measured (1 GiB, c2 vs c3, GB/s): mixed 4 KiB 75.7 / 72.5, 16 KiB 102.6 / 102.9, 32 KiB 106.4 / 107.9, 64 KiB 106.9 / 109.4;
        // run-length data 4 KiB 62.8 / 60.5, 16 KiB 97 / 131, 32 KiB 102 / 177, 64 KiB 104 / 215; random 64 KiB 158 / 176
        CU(c3_launch_compress(d_in, s.uoff, d_usz, s.stage, stride, d_csz, (u32)n, s.ticket, c->sm_count, st,
                              c->prof_on ? (u64*)c->prof.p : nullptr));
        c->launches++;
    } else if (c->compress_gen >= 2 && c1_compress_supported(max_chunk)) {
        CU(c2_launch_compress(d_in, s.uoff, d_usz, s.stage, stride, d_csz, (u32)n, s.ticket, c->sm_count, st, nullptr, nullptr, nullptr,
                              nullptr, c->prof_on ? (u64*)c->prof.p : nullptr, (u32)c->accel, c->c1mode));
        c->launches++;
    } else if (c->compress_gen == 1 && c1_compress_supported(max_chunk)) {
        CU(c1_launch_compress(d_in, s.uoff, d_usz, s.stage, stride, d_csz, (u32)n, s.ticket,
                              c->prof_on ? (u64*)c->prof.p : nullptr, c->sm_count, max_chunk, st, c->c1mode));
        c->launches++;
    } else {
        const size_t smem = (size_t)V0_WARPS * V0_TABLE * sizeof(u16);
        CU(cudaFuncSetAttribute(compress_v0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: every launch
        compress_v0_kernel<<<(unsigned)((n + V0_WARPS - 1) / V0_WARPS), V0_WARPS * 32, smem, st>>>(
            d_in, s.uoff, d_usz, s.stage, stride, d_csz, (u32)n);
        c->launches++;
    }
    time_end(c, st);
    const unsigned gn = (unsigned)((n + 255) / 256);
    if (frame) {
        frame_sizes_kernel<<<gn, 256, 0, st>>>(d_csz, d_usz, s.fsz, (u32)n, frame == 2 ? 4u : 0u);
        c->launches++;
    }
    if (wait_before_scan) CU(cudaStreamWaitEvent(st, wait_before_scan, 0));
    ScanJob out_scan{ frame ? s.fsz : d_csz, s.coff, d_total, base, mirror };
    scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>(out_scan, none, (u64)n);
    c->launches++;
    if (record_after_scan) CU(cudaEventRecord(record_after_scan, st));
    if (frame) {
        frame_headers_kernel<<<gn, 256, 0, st>>>(d_out, s.coff, d_csz, d_usz, (u32)n);
        c->launches++;
    }
    if (seg) {
        const size_t jobs_max = n * segs_per_chunk(max_chunk);
        const unsigned grid = (unsigned)std::min<size_t>(jobs_max, (size_t)c->sm_count * 16);
        seg_stitch_copy_kernel<<<grid, SEG_COPY_THREADS, 0, st>>>(sa, s.stage, slot_stride_for(SEG_BYTES), d_in, s.uoff, d_usz,
                                                                   d_csz, s.coff, d_out, (u32)jobs_max);
    } else {
        const unsigned grid = (unsigned)std::min<size_t>(n, (size_t)c->sm_count * 16);
        compact_kernel<<<grid, COMPACT_THREADS, 0, st>>>(s.stage, stride, d_in, s.uoff, d_usz, d_csz, s.coff, d_out, (u32)n);
    }
    c->launches++;
    if (frame == 2) {
        frame_block_xxh_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(d_out, s.coff, d_csz, (u32)n, 0, nullptr);
        c->launches++;
    }
    CU(cudaGetLastError());
    return 0;
}

int decompress_slice(lz4b200_ctx* c, const u8* d_in, const u64* d_csz, const u64* d_usz, size_t n, size_t max_chunk,
                     u8* d_out, u32* d_bad, const Scratch& s_in, cudaStream_t st, int team = 0,
                     const u64* given_coff = nullptr, u32 cap_all = 0)
{
    // given_coff: the chunks are not contiguous in d_in (frame blocks have a size word between them); the caller has
    // filled their positions and s.uoff (absolute output offsets) itself.
    Scratch s = s_in;
    if (given_coff) s.coff = const_cast<u64*>(given_coff);
    else {
        ScanJob a{ d_csz, s.coff, nullptr, nullptr }, b{ d_usz, s.uoff, nullptr, nullptr };
        scan_sizes_kernel<<<2, SCAN_THREADS, 0, st>>>(a, b, (u64)n);
        c->launches++;
    }
    time_begin(c, st);
    // team width: with few chunks per launch the run time is the latency of one chunk's chain, which is shorter
    // on wider teams, and wider teams also spread few chunks over more SMs
    if (!team) team = c->team;
    if (!team) {
        // measured (scripts/team_probe.py, gpu_probe.py): 1024 chunks take 2.0 ms on 16 lanes, 2.6 on 8, 3.7 on a warp (v2);
        // 4-lane teams (8 chunks per warp) only pay once there are enough chunks to fill the machine with them
        // (262 144 chunks of 4 KiB: 357 vs 301 GB/s; 65 536 of 16 KiB: 250 vs 294)
        if (c->kernel_gen == 5 || c->kernel_gen == 7) team = n * 2 <= (size_t)c->sm_count * 32 ? 16 : (n >= (size_t)c->sm_count * 7 * 32 * 4 ? 4 : 8);
        else if (c->kernel_gen >= 5) team = n * 2 <= (size_t)c->sm_count * 32 ? 16 : 8;
        else team = n * 4 <= (size_t)c->sm_count * 32 ? 32 : (n * 2 <= (size_t)c->sm_count * 32 ? 16 : 8);
    }
    if (c->kernel_gen == 7 && s.v7_table && max_chunk <= V7_MAX_CHUNK) {
        // sequence table, lane-per-sequence decoder, then the team decoder for the chunks left on the list.
        // A large batch is cut in two: while the table of one part is being built -- a latency-bound pass that leaves the
        // machine almost idle (one lane per chunk) -- the team decoder, whose run time is the latency of a chunk's chain
        // rather than the number of chunks, works on the other part on a second stream.
        for (size_t b = 0; b < n; b += V7_SUB) {
            const size_t m = std::min(V7_SUB, n - b);
            const int pct = c->split >= 0 ? c->split : 0;       // measured: the table pass is a latency-bound chain, and ANY co-runner slows it (1 GiB: 4.1 ms alone, 6.0 ms at 20 %)
            const size_t mA = (m >= 8192 && c->aux_stream) ? m * (size_t)pct / 100 : 0;   // team decoder's share: chunks [b, b + mA)
            if (mA) {
                if (!c->ev_fork[s.slice]) {
                    CU(cudaEventCreateWithFlags(&c->ev_fork[s.slice], cudaEventDisableTiming));
                    CU(cudaEventCreateWithFlags(&c->ev_join[s.slice], cudaEventDisableTiming));
                }
                CU(cudaEventRecord(c->ev_fork[s.slice], st));
                CU(cudaStreamWaitEvent(c->aux_stream, c->ev_fork[s.slice], 0));
            }
            const size_t b2 = b + mA, m2 = m - mA;
            if (m2) {
                V7Params P{ d_in, s.coff + b2, d_csz + b2, s.uoff + b2, d_usz + b2, d_out, s.status + b2, d_bad, s.v7_count + 1, (u32)m2, cap_all,
                            s.v7_table, s.v7_cap, s.v7_nseq + b2, s.v7_route + b2, s.v7_list + b2, s.v7_count, s.v7_order + b2, s.v7_bucket, nullptr };
                CU(v7_launch_parse(P, st, c->v7_lanes));
                if (mA) {
                    CU(v5_launch_decompress(16, d_in, s.coff + b, d_csz + b, s.uoff + b, d_usz + b, d_out, s.status + b, d_bad, (u32)mA,
                                            s.ticket2, c->sm_count, c->aux_stream, cap_all));
                    CU(cudaEventRecord(c->ev_join[s.slice], c->aux_stream));
                    c->launches += 1;
                }
                CU(v7_launch_decompress(P, c->sm_count, st));
                CU(v5_launch_decompress(team == 32 ? 16 : team, d_in, s.coff + b2, d_csz + b2, s.uoff + b2, d_usz + b2, d_out,
                                        s.status + b2, d_bad, (u32)m2, s.v7_count + 2, c->sm_count, st, cap_all, s.v7_list + b2, s.v7_count, true));
                c->launches += 4;
            }
            if (mA) CU(cudaStreamWaitEvent(st, c->ev_join[s.slice], 0));
        }
    } else if (c->kernel_gen >= 7 && c->coarse_chunks && max_chunk > V7_MAX_CHUNK && s.first_chunk + n <= c->coarse_chunks) {
        // large chunks: the ones that are a handful of long copies are cut into 32 KiB pieces (kernels_coarse.cuh), the team
        // decoder takes the rest from a list
        u8* b = (u8*)c->coarse.p;
        const size_t nt = c->coarse_chunks;
        CoarseParams P{ d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, (u32)n, cap_all,
                        (uint4*)(b + 1024) + s.first_chunk * C_MAXD,
                        (uint2*)(b + 1024 + nt * C_MAXD * sizeof(uint4)) + s.first_chunk * C_MAXD,
                        (u32*)(b + 1024 + nt * C_MAXD * (sizeof(uint4) + sizeof(uint2))) + s.first_chunk,
                        (u32*)(b + 1024 + nt * C_MAXD * (sizeof(uint4) + sizeof(uint2))) + nt + s.first_chunk,
                        (u32*)b + s.slice, (u32*)b + 64 + 2 * s.slice };
        // the copy kernel synchronises its CTAs through a counter, so all of them must be resident: two of them on two
        // streams (pipeline slices) could each hold the slots the other waits for -- one at a time per context
        if (!c->ev_coarse) CU(cudaEventCreateWithFlags(&c->ev_coarse, cudaEventDisableTiming));
        else CU(cudaStreamWaitEvent(st, c->ev_coarse, 0));
        const cudaError_t ce = coarse_launch(P, c->sm_count, st);
        if (ce != cudaSuccess) {
            // (a cooperative launch the device cannot hold: the descriptors are dropped, the team decoder takes every chunk,
            // and this context stops trying)
            cudaGetLastError();
            c->coarse_on = false;
            CU(v5_launch_decompress(team == 32 ? 16 : team, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                    s.ticket, c->sm_count, st, cap_all));
            c->launches += 3;
        } else {
        CU(cudaEventRecord(c->ev_coarse, st));
        if (s.v7_table && n <= V7_SUB && max_chunk <= V7_BIG_MAX) {
            // ... and the rest (compressible data) goes through the sequence table as 64 KiB chunks do: one lane per chunk
            // walks the chain (the entries keep their 16-bit fields, pass 2 counts the wrap-arounds), one lane per sequence
            // copies; what the table cannot hold is left to the team decoder
            V7Params Q{ d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, s.v7_count + 1, (u32)n, cap_all,
                        s.v7_table, s.v7_cap, s.v7_nseq, s.v7_route, s.v7_list, s.v7_count, s.v7_order, s.v7_bucket, P.ndesc };
            CU(v7_launch_parse(Q, st, 32, true));
            CU(v7_launch_decompress(Q, c->sm_count, st, true));
            CU(v5_launch_decompress(team == 32 ? 16 : team, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                    s.v7_count + 2, c->sm_count, st, cap_all, s.v7_list, s.v7_count, true));
            c->launches += 7;
        } else {
            CU(v5_launch_decompress(team == 32 ? 16 : team, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                    s.ticket, c->sm_count, st, cap_all, P.list, P.list_count));
            c->launches += 3;
        }
        }
    } else if (c->kernel_gen >= 7) {
        CU(v5_launch_decompress(team == 32 ? 16 : team, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                s.ticket, c->sm_count, st, cap_all));
        c->launches += 1;
    } else if (c->kernel_gen == 6 && team != 32) {
        CU(v6_launch_decompress(team, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                s.ticket, c->sm_count, st, c->v6depth));
        c->launches += 1;
    } else if (c->kernel_gen == 5 && team != 32) {
        CU(v5_launch_decompress(team, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                s.ticket, c->sm_count, st, cap_all));
        c->launches += 1;
    } else if (c->kernel_gen == 4 && team != 32) {
        CU(v4_launch_decompress(team, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                s.ticket, c->sm_count, st));
        c->launches += 1;
    } else if (c->kernel_gen >= 4) {                    // one warp per chunk: the v2 kernel at T = 32
        CU(v2_launch_decompress(32, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                s.ticket, c->sm_count, st));
        c->launches += 1;
    } else if (c->kernel_gen == 3) {
        CU(v3_launch_decompress(team == 32 ? 16 : team, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                s.ticket, c->sm_count, st));
        c->launches += 1;
    } else if (c->kernel_gen == 2) {
        CU(v2_launch_decompress(team, d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n,
                                s.ticket, c->sm_count, st));
        c->launches += 1;
    } else if (c->kernel_gen == 1 && v1_decompress_supported(max_chunk)) {
        CU(v1_launch_decompress(d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n, s.ticket,
                                c->prof_on ? (u64*)c->prof.p : nullptr, c->sm_count, max_chunk, st));
        c->launches += 1;
    } else {
        decompress_v0_kernel<<<(unsigned)((n + V0_WARPS - 1) / V0_WARPS), V0_WARPS * 32, 0, st>>>(
            d_in, s.coff, d_csz, s.uoff, d_usz, d_out, s.status, d_bad, (u32)n);
        c->launches++;
    }
    time_end(c, st);
    CU(cudaGetLastError());
    return 0;
}

// lzbench allocates inbuf/compbuf/decomp with malloc (_lzbench/lzbench.cpp:256-263) and reuses them for every
// iteration: page-locking them once turns every later cudaMemcpyAsync into a direct DMA.  Best effort.
void maybe_register(lz4b200_ctx* c, const char* p, size_t n)
{
    if (!c->host_register || n < (1u << 20)) return;
    for (int i = 0; i < c->n_regs; i++)
        if (p >= c->regs[i].p && p + n <= c->regs[i].p + c->regs[i].n) return;
    if (c->n_regs == 16) {                               // drop the oldest registration
        cudaHostUnregister(const_cast<char*>(c->regs[0].p)); cudaGetLastError();
        for (int i = 1; i < 16; i++) c->regs[i - 1] = c->regs[i];
        c->n_regs = 15;
    }
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered) return;
    cudaGetLastError();
    if (cudaHostRegister(const_cast<char*>(p), n, cudaHostRegisterDefault) == cudaSuccess) c->regs[c->n_regs++] = { p, n };
    else cudaGetLastError();
}

int check_batch(size_t n, size_t max_chunk)
{
    if (n > 0x7FFFFFFFull) return fail(-1, "too many chunks");
    if (max_chunk > 0x7E000000ull) return fail(-1, "chunk larger than LZ4_MAX_INPUT_SIZE");
    return 0;
}

// Cut chunks [0,n) into K contiguous slices of roughly equal uncompressed size; first[k] = first chunk of slice k.
// Equal slices by bytes.  Slice counts measured with scripts/e2e_probe.py (1 GiB, pinned buffers, 8 streams): both
// directions are best with 16 slices (the PCIe copy of the larger side is the bottleneck as soon as the first slice
// is through, and a slice's kernels last as long as one chunk's serial chain however few chunks it holds, so
// several slices must be in flight at once).  Decompression starts with three smaller slices (see below).
int plan_slices(const lz4b200_ctx* c, const size_t* chunk_sizes, size_t n, size_t total, bool decompress,
                size_t* first /* K+1 */)
{
    int K = c->slices;
    if (K <= 0) {
        K = total >= (512u << 20) ? 16 : (total >= (128u << 20) ? 8 : (total >= (32u << 20) ? 2 : 1));
    }
    if (K > lz4b200_ctx::MAX_SLICES) K = lz4b200_ctx::MAX_SLICES;
    if ((size_t)K > n) K = (int)n;
    // boundaries as cumulative fractions of the bytes: equal slices, or (option "ramp") a first slice of 1/8 of the
    // regular size that doubles up to it, so that the outbound copy starts earlier
    double cut[lz4b200_ctx::MAX_SLICES + 1];
    for (int k = 0; k <= K; k++) cut[k] = (double)k / K;
    if (c->ramp && decompress && K >= 8) {      // measured (scripts/h2d_ab.py): decompress 24.1 vs 24.7 ms per GiB, compress no gain
        const double r[3] = { 1.0 / 8, 2.0 / 8, 4.0 / 8 };
        const double reg = 1.0 / (K - 3 + r[0] + r[1] + r[2]);      // regular slice size so that everything sums to 1
        double acc_r = 0;
        for (int k = 0; k < K; k++) { acc_r += (k < 3 ? r[k] : 1.0) * reg; cut[k + 1] = acc_r; }
        cut[K] = 1.0;
    }
    size_t acc = 0, k = 1;
    first[0] = 0;
    for (size_t i = 0; i < n && k < (size_t)K; i++) {
        acc += chunk_sizes[i];
        if ((double)acc >= (double)total * cut[k] && i + 1 < n) first[k++] = i + 1;
    }
    K = (int)k;
    first[K] = n;
    return K;
}

int ensure_events(lz4b200_ctx* c, int K)
{
    if (!c->ev_sizes) CU(cudaEventCreateWithFlags(&c->ev_sizes, cudaEventDisableTiming));
    for (int k = 0; k < K; k++) {
        if (!c->ev_scan[k]) CU(cudaEventCreateWithFlags(&c->ev_scan[k], cudaEventDisableTiming));
        if (!c->ev_done[k]) CU(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
        if (!c->ev_in[k]) CU(cudaEventCreateWithFlags(&c->ev_in[k], cudaEventDisableTiming));
    }
    return 0;
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------

extern "C" {

const char* lz4b200_version(void) { return LZ4B200_VERSION_STRING; }
const char* lz4b200_last_error(void) { return g_err; }

size_t lz4b200_compress_bound(size_t n)
{
    if (n > 0x7E000000ull) return 0;
    return n + n / 255 + 16;
}

int lz4b200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(-2, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int lz4b200_create(int device, size_t max_total_bytes, size_t max_chunk_bytes, lz4b200_ctx** out)
{
    if (!out) return fail(-1, "null ctx pointer");
    *out = nullptr;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(-2, "no CUDA device: lz4b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(-1, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10 || prop.minor != 0) return fail(-2, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    lz4b200_ctx* c = new (std::nothrow) lz4b200_ctx();
    if (!c) return fail(-3, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (const char* gen = getenv("LZ4B200_KERNELS")) lz4b200_set_option(c, "kernels", gen);   // same meaning as the option
    for (int i = 0; i < lz4b200_ctx::NS; i++) {
        cudaError_t e = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) { lz4b200_destroy(c); return fail(-2, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    }
    if (cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking) != cudaSuccess) {
        lz4b200_destroy(c); return fail(-2, "cudaStreamCreate failed");
    }
    if (cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess) {
        lz4b200_destroy(c); return fail(-2, "cudaStreamCreate failed");
    }
    if (cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) {
        lz4b200_destroy(c); return fail(-2, "cudaStreamCreate failed");
    }
    if (max_total_bytes && max_chunk_bytes) {
        size_t n = (max_total_bytes + max_chunk_bytes - 1) / max_chunk_bytes;
        // best-effort pre-sizing; growth on demand covers the rest
        c->d_plain.ensure(max_total_bytes + 16);
        c->d_packed.ensure(max_total_bytes + max_total_bytes / 255 + 16 * n + 16);
        c->stage.ensure(n * slot_stride_for(max_chunk_bytes));
        c->usz.ensure(n * 8); c->csz.ensure(n * 8); c->uoff.ensure(n * 8); c->coff.ensure(n * 8);
        c->status.ensure(n * 4);
        c->h_sizes.ensure((2 * n + lz4b200_ctx::MAX_SLICES + 2) * 8);
        cudaGetLastError();                              // best effort: a failed allocation here must not surface at the first launch
    }
    if (ensure_scratch(c, 1, 0) != 0) { lz4b200_destroy(c); return -2; }
    if (cudaHostAlloc((void**)&c->h_mirror, lz4b200_ctx::MAX_SLICES * sizeof(u64), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&c->d_mirror, c->h_mirror, 0) != cudaSuccess) {
        lz4b200_destroy(c); return fail(-2, "cudaHostAlloc (mapped) failed");
    }
    *out = c;
    return 0;
}

void lz4b200_destroy(lz4b200_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < c->n_regs; i++) cudaHostUnregister(const_cast<char*>(c->regs[i].p));
    for (int i = 0; i < lz4b200_ctx::NS; i++) if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->ev_sizes) cudaEventDestroy(c->ev_sizes);
    if (c->ev_coarse) cudaEventDestroy(c->ev_coarse);
    for (int i = 0; i < lz4b200_ctx::MAX_SLICES; i++) {
        if (c->ev_scan[i]) cudaEventDestroy(c->ev_scan[i]);
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_fork[i]) cudaEventDestroy(c->ev_fork[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    c->usz.release(); c->csz.release(); c->uoff.release(); c->coff.release();
    c->scalars.release(); c->status.release(); c->stage.release(); c->seg64.release(); c->seg32.release();
    c->d_plain.release(); c->d_packed.release(); c->h_sizes.release(); c->prof.release();
    c->fsz.release(); c->fhdr.release(); c->v7_table.release(); c->v7_u32.release();
    if (c->h_mirror) cudaFreeHost(c->h_mirror);
    for (int i = 0; i < lz4b200_ctx::MAX_TIMED; i++) if (c->ev0[i]) { cudaEventDestroy(c->ev0[i]); cudaEventDestroy(c->ev1[i]); }
    delete c;
}

uint64_t lz4b200_launch_count(const lz4b200_ctx* c) { return c ? c->launches : 0; }

size_t lz4b200_compress_device_bound(size_t n_chunks, size_t max_chunk)
{
    // compact image: every chunk at most its bound; the segmented path (chunks > 64 KiB) adds up to 16 bytes per 48 KiB segment
    const size_t b = lz4b200_compress_bound(max_chunk);
    if (!b) return 0;
    return n_chunks * (b + 16 * (max_chunk / SEG_NEW + 2)) + 16;
}

int lz4b200_unregister_host(lz4b200_ctx* c, const void* p)
{
    if (!c || !p) return -1;
    for (int i = 0; i < c->n_regs; i++) {
        const char* q = (const char*)p;
        if (q >= c->regs[i].p && q < c->regs[i].p + c->regs[i].n) {
            cudaSetDevice(c->device);
            cudaHostUnregister(const_cast<char*>(c->regs[i].p)); cudaGetLastError();
            for (int j = i + 1; j < c->n_regs; j++) c->regs[j - 1] = c->regs[j];
            c->n_regs--;
            return 0;
        }
    }
    return -1;
}

int lz4b200_set_option(lz4b200_ctx* c, const char* key, const char* value)
{
    if (!c || !key || !value) return -1;
    if (!strcmp(key, "kernels")) {          // decode generation (and v0 also selects the v0 compressor)
        if (!strcmp(value, "v0")) { c->kernel_gen = 0; c->compress_gen = 0; return 0; }
        if (!strcmp(value, "v1")) { c->kernel_gen = 1; c->compress_gen = 1; return 0; }      // first tiled generation of both
        if (!strcmp(value, "v2")) { c->kernel_gen = 2; c->compress_gen = 2; return 0; }
        if (!strcmp(value, "v3")) { c->kernel_gen = 3; c->compress_gen = 2; return 0; }
        if (!strcmp(value, "v4")) { c->kernel_gen = 4; c->compress_gen = 2; return 0; }
        if (!strcmp(value, "v5")) { c->kernel_gen = 5; c->compress_gen = 2; return 0; }
        if (!strcmp(value, "v6")) { c->kernel_gen = 6; c->compress_gen = 2; return 0; }
        if (!strcmp(value, "v7") || !strcmp(value, "default")) {
            c->kernel_gen = 7; c->compress_gen = strcmp(value, "v7") ? 3 : 2; c->c3_forced = false; c->v7_min_chunk = strcmp(value, "v7") ? 16385 : 0;
            return 0;
        }
    }
    if (!strcmp(key, "compress")) {
        if (!strcmp(value, "v0")) { c->compress_gen = 0; c->c3_forced = false; return 0; }
        if (!strcmp(value, "v1")) { c->compress_gen = 1; c->c3_forced = false; return 0; }
        if (!strcmp(value, "v2")) { c->compress_gen = 2; c->c3_forced = false; return 0; }
        if (!strcmp(value, "v3")) { c->compress_gen = 3; c->c3_forced = true; return 0; }
        if (!strcmp(value, "default")) { c->compress_gen = 3; c->c3_forced = false; return 0; }
    }
    if (!strcmp(key, "c1mode")) { c->c1mode = atoi(value); return 0; }
    if (!strcmp(key, "accel")) {                        // LZ4_compress_fast's acceleration (lz4/lz4.c:1352-1375); the lz4fast levels of lzbench
        int a = atoi(value);
        c->accel = a < 1 ? 1 : (a > 65537 ? 65537 : a);
        return 0;
    }
    if (!strcmp(key, "team")) {
        int t = atoi(value);
        if (t == 0 || t == 4 || t == 8 || t == 16 || t == 32) { c->team = t; return 0; }
    }
    if (!strcmp(key, "depth")) {                        // v6: iterations between fetching and storing a match part
        int d = atoi(value);
        if (d >= 2 && d <= 4) { c->v6depth = d; return 0; }
    }
    if (!strcmp(key, "split")) {                        // v7: percent of a large batch decoded by the team decoder beside the table pass (-1 = automatic)
        int p = atoi(value);
        if (p >= -1 && p <= 90) { c->split = p; return 0; }
    }
    if (!strcmp(key, "coarse")) { c->coarse_on = value[0] == '1'; return 0; }
    if (!strcmp(key, "v7big")) { c->v7_big = value[0] == '1'; return 0; }
    if (!strcmp(key, "planes")) { const int v = atoi(value); if (v == 8 || v == 16 || v == 32) { c->v7_lanes = v; return 0; } return -1; }
    if (!strcmp(key, "ramp")) { c->ramp = value[0] == '1'; return 0; }
    if (!strcmp(key, "h2d")) { c->own_h2d = strcmp(value, "shared") != 0; return 0; }
    if (!strcmp(key, "host_register")) { c->host_register = value[0] == '1'; return 0; }
    if (!strcmp(key, "slices")) {
        int k = atoi(value);
        if (k >= 0 && k <= lz4b200_ctx::MAX_SLICES) { c->slices = k; return 0; }
    }
    if (!strcmp(key, "prof")) {
        c->prof_on = value[0] == '1';
        if (c->prof_on) { if (c->prof.ensure(128) != cudaSuccess) return fail(-2, "prof alloc"); cudaMemset(c->prof.p, 0, 128); }
        return 0;
    }
    if (!strcmp(key, "timing")) { c->timing = value[0] == '1'; c->n_timed = 0; return 0; }
    return fail(-1, "unknown option %s=%s", key, value);
}

int lz4b200_debug_prof(lz4b200_ctx* c, uint64_t* out16)
{
    if (!c || !c->prof.p) return fail(-1, "prof not enabled");
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out16, c->prof.p, 128, cudaMemcpyDeviceToHost));
    CU(cudaMemset(c->prof.p, 0, 128));
    return 0;
}

int lz4b200_kernel_times_ms(lz4b200_ctx* c, float* out_ms, int max_n)
{
    if (!c) return fail(-1, "null ctx");
    int n = std::min(c->n_timed, max_n);
    for (int i = 0; i < n; i++) {
        CU(cudaEventSynchronize(c->ev1[i]));
        CU(cudaEventElapsedTime(&out_ms[i], c->ev0[i], c->ev1[i]));
    }
    c->n_timed = 0;
    return n;
}

int lz4b200_compress_device(lz4b200_ctx* c, const void* d_in, const uint64_t* d_chunk_sizes, size_t n,
                            size_t max_chunk, void* d_out, uint64_t* d_compr_sizes,
                            uint64_t* d_total_out, void* stream)
{
    if (!c) return fail(-1, "null ctx");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->device));
    if (n == 0) { if (d_total_out) CU(cudaMemsetAsync(d_total_out, 0, sizeof(u64), st)); return 0; }
    if (int rc = check_batch(n, max_chunk)) return rc;
    if (int rc = ensure_scratch(c, n, stage_bytes_for(c, n, max_chunk))) return rc;
    if (segmented(c, max_chunk)) { if (int rc = ensure_segmented(c, n, max_chunk)) return rc; c->seg_chunks = n; }
    u64* total = d_total_out ? (u64*)d_total_out : (u64*)((u8*)c->scalars.p + SC_TOTAL);
    return compress_slice(c, (const u8*)d_in, (const u64*)d_chunk_sizes, n, max_chunk, (u8*)d_out,
                          (u64*)d_compr_sizes, nullptr, total, scratch_at(c, 0, stage_bytes_for(c, 1, max_chunk), 0), st,
                          nullptr, nullptr);
}

int lz4b200_decompress_device(lz4b200_ctx* c, const void* d_in, const uint64_t* d_compr_sizes,
                              const uint64_t* d_chunk_sizes, size_t n, size_t max_chunk,
                              void* d_out, uint32_t* d_bad_chunks, void* stream)
{
    if (!c) return fail(-1, "null ctx");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->device));
    if (int rc = ensure_scratch(c, n ? n : 1, 0)) return rc;
    ensure_v7(c, n, std::min(n, V7_SUB), max_chunk);
    ensure_coarse(c, n, max_chunk);
    u32* bad = d_bad_chunks ? d_bad_chunks : (u32*)((u8*)c->scalars.p + SC_BAD);
    CU(cudaMemsetAsync(bad, 0, sizeof(u32), st));
    if (n == 0) return 0;
    if (int rc = check_batch(n, max_chunk)) return rc;
    return decompress_slice(c, (const u8*)d_in, (const u64*)d_compr_sizes, (const u64*)d_chunk_sizes, n, max_chunk,
                            (u8*)d_out, bad, scratch_at(c, 0, 0, 0), st);
}

// Host-buffer compress: slices of the chunk list flow through H2D -> kernels -> D2H on NS streams so the
// PCIe copies of one slice overlap the kernels of another.  The compacted position of a slice depends on
// the compressed size of the slices before it, which exists only on the device: the output scan of slice
// k starts from the running total slice k-1 left in device memory (event-ordered), and the host learns
// the totals one slice behind to place the D2H copies.
// frame_hdr > 0: write LZ4 frame blocks starting at out + frame_hdr instead of lzbench's contiguous image, and keep
// frame_tail bytes of out_capacity free behind them (end mark, content checksum); the caller writes header and tail.
static int64_t compress_batch_impl(lz4b200_ctx* c, const char* in, const size_t* chunk_sizes, size_t n,
                                   char* out, size_t out_capacity, size_t* compr_sizes,
                                   size_t frame_hdr, size_t frame_tail, bool block_checksum)
{
    if (!c) return fail(-1, "null ctx");
    if (n == 0) return 0;
    if (!in || !chunk_sizes || !out || !compr_sizes) return fail(-1, "null argument");
    CU(cudaSetDevice(c->device));
    size_t total = 0, max_chunk = 0;
    CU(c->h_sizes.ensure((2 * n + lz4b200_ctx::MAX_SLICES + 2) * sizeof(u64)));
    u64* h = (u64*)c->h_sizes.p;
    for (size_t i = 0; i < n; i++) { h[i] = chunk_sizes[i]; total += chunk_sizes[i]; max_chunk = std::max(max_chunk, chunk_sizes[i]); }
    if (int rc = check_batch(n, max_chunk)) return rc;
    maybe_register(c, in, total);
    maybe_register(c, out, out_capacity);
    const size_t stride = stage_bytes_for(c, 1, max_chunk);          // staging bytes per chunk
    CU(c->usz.ensure(n * 8)); CU(c->csz.ensure(n * 8));
    CU(c->d_plain.ensure(total + 16));
    CU(c->d_packed.ensure(total + total / 255 + 16 * (n + total / SEG_BYTES + 1) + 16 + (frame_hdr ? 8 * n + 32 : 0)));
    if (frame_hdr) CU(c->fsz.ensure(n * 8));
    if (int rc = ensure_scratch(c, n, n * stride)) return rc;
    if (segmented(c, max_chunk)) { if (int rc = ensure_segmented(c, n, max_chunk)) return rc; c->seg_chunks = n; }
    size_t first[lz4b200_ctx::MAX_SLICES + 1];
    const int K = plan_slices(c, chunk_sizes, n, total, false, first);
    if (int rc = ensure_events(c, K)) return rc;

    u64* running = (u64*)((u8*)c->scalars.p + SC_RUNNING);
    u64* fbase = (u64*)((u8*)c->scalars.p + SC_FBASE);
    const int frame = frame_hdr ? (block_checksum ? 2 : 1) : 0;
    if (frame) {                                          // header is at most 19 bytes: the low byte carries it
        CU(cudaMemsetAsync(fbase, 0, 8, c->streams[0]));
        CU(cudaMemsetAsync(fbase, (int)frame_hdr, 1, c->streams[0]));
    }
    const size_t usable = out_capacity >= frame_tail ? out_capacity - frame_tail : 0;
    CU(cudaMemcpyAsync(c->usz.p, h, n * 8, cudaMemcpyHostToDevice, c->streams[0]));
    CU(cudaEventRecord(c->ev_sizes, c->streams[0]));
    size_t in_off = 0;
    for (int k = 0; k < K; k++) {
        cudaStream_t st = c->streams[k % lz4b200_ctx::NS];
        const size_t f = first[k], m = first[k + 1] - first[k];
        size_t bytes = 0;
        for (size_t i = f; i < f + m; i++) bytes += chunk_sizes[i];
        CU(cudaStreamWaitEvent(st, c->ev_sizes, 0));
        // inbound copies on their own stream, except for frames: measured on one box (scripts/h2d_ab.py, 1 GiB, six runs)
        // chunked compress 24.1-24.7 ms own vs 24.4-25.5 shared, frame compress 24.6-30.6 (erratic) own vs 24.4-25.8 shared
        const bool own = c->own_h2d && !frame;
        CU(cudaMemcpyAsync((u8*)c->d_plain.p + in_off, in + in_off, bytes, cudaMemcpyHostToDevice, own ? c->h2d_stream : st));
        if (own) { CU(cudaEventRecord(c->ev_in[k], c->h2d_stream)); CU(cudaStreamWaitEvent(st, c->ev_in[k], 0)); }
        int rc = compress_slice(c, (const u8*)c->d_plain.p + in_off, (const u64*)c->usz.p + f, m, max_chunk,
                                (u8*)c->d_packed.p, (u64*)c->csz.p + f, k ? running + (k - 1) : (frame ? fbase : nullptr), running + k,
                                scratch_at(c, f, stride, k), st, k ? c->ev_scan[k - 1] : nullptr, c->ev_scan[k], frame, c->d_mirror + k);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        // no device-to-host copies on this stream: they would queue behind the bulk copies of earlier slices on the
        // copy engine and hold back this stream's next H2D; totals reach the host through the mapped mirror
        CU(cudaEventRecord(c->ev_done[k], st));
        in_off += bytes;
    }
    u64 base = frame_hdr;
    bool overflow = false;
    for (int k = 0; k < K; k++) {
        CU(cudaEventSynchronize(c->ev_done[k]));
        const u64 end = ((volatile u64*)c->h_mirror)[k];
        if (end > usable) { overflow = true; break; }
        CU(cudaMemcpyAsync(out + base, (u8*)c->d_packed.p + base, end - base, cudaMemcpyDeviceToHost, c->d2h_stream));
        base = end;
    }
    if (overflow) {
        CU(cudaDeviceSynchronize());
        fail(0, "compressed image exceeds output capacity %zu", out_capacity);
        return 0;                                        // _lzbench/lzbench.cpp:286
    }
    CU(cudaMemcpyAsync(h + n, c->csz.p, n * 8, cudaMemcpyDeviceToHost, c->d2h_stream));
    CU(cudaStreamSynchronize(c->d2h_stream));
    for (size_t i = 0; i < n; i++) compr_sizes[i] = (size_t)h[n + i];
    return (int64_t)base;
}

int64_t lz4b200_compress_batch(lz4b200_ctx* c, const char* in, const size_t* chunk_sizes, size_t n,
                               char* out, size_t out_capacity, size_t* compr_sizes)
{
    return compress_batch_impl(c, in, chunk_sizes, n, out, out_capacity, compr_sizes, 0, 0, false);
}

int64_t lz4b200_decompress_batch(lz4b200_ctx* c, const char* in, const size_t* compr_sizes,
                                 const size_t* chunk_sizes, size_t n, char* out)
{
    if (!c) return fail(-1, "null ctx");
    if (n == 0) return 0;
    if (!in || !chunk_sizes || !out || !compr_sizes) return fail(-1, "null argument");
    CU(cudaSetDevice(c->device));
    size_t total = 0, ctotal = 0, max_chunk = 0;
    CU(c->h_sizes.ensure((2 * n + lz4b200_ctx::MAX_SLICES + 2) * sizeof(u64)));
    u64* h = (u64*)c->h_sizes.p;
    for (size_t i = 0; i < n; i++) {
        h[i] = chunk_sizes[i]; h[n + i] = compr_sizes[i];
        total += chunk_sizes[i]; ctotal += compr_sizes[i];
        max_chunk = std::max(max_chunk, chunk_sizes[i]);
    }
    if (int rc = check_batch(n, max_chunk)) return rc;
    maybe_register(c, in, ctotal);
    maybe_register(c, out, total);
    CU(c->usz.ensure(n * 8)); CU(c->csz.ensure(n * 8));
    CU(c->d_plain.ensure(total + 16));
    CU(c->d_packed.ensure(ctotal + 16));
    if (int rc = ensure_scratch(c, n, 0)) return rc;
    ensure_v7(c, n, n, max_chunk);
    ensure_coarse(c, n, max_chunk);
    size_t first[lz4b200_ctx::MAX_SLICES + 1];
    const int K = plan_slices(c, chunk_sizes, n, total, true, first);
    if (int rc = ensure_events(c, K)) return rc;

    u32* d_bad = (u32*)((u8*)c->scalars.p + SC_BAD);
    CU(cudaMemsetAsync(d_bad, 0, 4, c->streams[0]));
    CU(cudaMemcpyAsync(c->usz.p, h, n * 8, cudaMemcpyHostToDevice, c->streams[0]));
    CU(cudaMemcpyAsync(c->csz.p, h + n, n * 8, cudaMemcpyHostToDevice, c->streams[0]));
    CU(cudaEventRecord(c->ev_sizes, c->streams[0]));
    size_t uoff = 0, coff = 0;
    for (int k = 0; k < K; k++) {
        cudaStream_t st = c->streams[k % lz4b200_ctx::NS];
        const size_t f = first[k], m = first[k + 1] - first[k];
        size_t ub = 0, cb = 0;
        for (size_t i = f; i < f + m; i++) { ub += chunk_sizes[i]; cb += compr_sizes[i]; }
        CU(cudaStreamWaitEvent(st, c->ev_sizes, 0));
        CU(cudaMemcpyAsync((u8*)c->d_packed.p + coff, in + coff, cb, cudaMemcpyHostToDevice, c->own_h2d ? c->h2d_stream : st));
        if (c->own_h2d) { CU(cudaEventRecord(c->ev_in[k], c->h2d_stream)); CU(cudaStreamWaitEvent(st, c->ev_in[k], 0)); }
        int rc = decompress_slice(c, (const u8*)c->d_packed.p + coff, (const u64*)c->csz.p + f, (const u64*)c->usz.p + f,
                                  m, max_chunk, (u8*)c->d_plain.p + uoff, d_bad, scratch_at(c, f, 0, k), st);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        CU(cudaEventRecord(c->ev_done[k], st));
        CU(cudaStreamWaitEvent(c->d2h_stream, c->ev_done[k], 0));
        CU(cudaMemcpyAsync(out + uoff, (u8*)c->d_plain.p + uoff, ub, cudaMemcpyDeviceToHost, c->d2h_stream));
        uoff += ub; coff += cb;
    }
    u32* h_bad = (u32*)(h + 2 * n);
    CU(cudaMemcpyAsync(h_bad, d_bad, 4, cudaMemcpyDeviceToHost, c->d2h_stream));
    CU(cudaStreamSynchronize(c->d2h_stream));
    if (*h_bad) { fail(-1, "%u malformed chunk(s)", *h_bad); return -(int64_t)*h_bad; }
    return (int64_t)total;
}

// The yardstick for e2e: what the reference's `cudaMemcpy` row measures (lzbench_cuda_memcpy,
// _lzbench/compressors.cpp:1828-1833: H2D then D2H of the same bytes), through this library's plumbing -- page-locked
// caller buffers, slices on the pipeline streams so the two directions overlap.  No kernel runs.
int64_t lz4b200_memcpy_batch(lz4b200_ctx* c, const char* in, size_t n, char* out)
{
    if (!c) return fail(-1, "null ctx");
    if (n == 0) return 0;
    if (!in || !out) return fail(-1, "null argument");
    CU(cudaSetDevice(c->device));
    maybe_register(c, in, n);
    maybe_register(c, out, n);
    CU(c->d_plain.ensure(n + 16));
    const int K = n >= (512u << 20) ? 16 : (n >= (32u << 20) ? 4 : 1);
    if (int rc = ensure_events(c, K)) return rc;
    for (int k = 0; k < K; k++) {
        const size_t lo = n * k / K, hi = n * (k + 1) / K;
        cudaStream_t st = c->h2d_stream;
        CU(cudaMemcpyAsync((u8*)c->d_plain.p + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(c->ev_done[k], st));
        CU(cudaStreamWaitEvent(c->d2h_stream, c->ev_done[k], 0));
        CU(cudaMemcpyAsync(out + lo, (u8*)c->d_plain.p + lo, hi - lo, cudaMemcpyDeviceToHost, c->d2h_stream));
    }
    CU(cudaStreamSynchronize(c->d2h_stream));
    return (int64_t)n;
}

// ---- LZ4 frame container (frame.cuh) -----------------------------------------------------------

size_t lz4b200_frame_bound(size_t n, int block_id)
{
    const size_t b = frame_block_bytes(block_id);
    if (!b) return 0;
    return 19 + (n / b + 1) * 8 + n + 8;                   // header, per block size word + checksum, end mark, content checksum
}

// One frame, independent blocks of frame_block_bytes(block_id): what LZ4F_compressFrame produces with
// blockMode = LZ4F_blockIndependent (lz4/lz4frame.c:429-470), except that every block is compressed with a fresh
// table (as lzbench's chunks are) and the block size id is used as given.
int64_t lz4b200_frame_compress(lz4b200_ctx* c, const char* in, size_t n, char* out, size_t out_capacity,
                               int block_id, int flags)
{
    if (!c) return fail(-1, "null ctx");
    const size_t bsz = frame_block_bytes(block_id);
    if (!bsz) return fail(-1, "block_id must be 4..7 (64 KiB .. 4 MiB)");
    if (!out || (n && !in)) return fail(-1, "null argument");
    const bool cck = flags & LZ4B200_FRAME_CONTENT_CHECKSUM, bck = flags & LZ4B200_FRAME_BLOCK_CHECKSUM,
               wsz = flags & LZ4B200_FRAME_CONTENT_SIZE;
    const size_t tail = 4 + (cck ? 4 : 0);
    if (out_capacity < 19 + tail) return fail(-1, "output capacity too small for a frame");
    u8 hdr[19];
    const size_t hl = frame_write_header(hdr, block_id, bck, cck, wsz, n);
    int64_t end = (int64_t)hl;
    if (n) {
        const size_t nb = (n + bsz - 1) / bsz;
        std::vector<size_t> sizes(nb, bsz), csz(nb);
        sizes[nb - 1] = n - (nb - 1) * bsz;
        // the block pipeline runs asynchronously up to its first host wait; the content checksum (one serial
        // xxHash32 chain over the input, host only) would ideally overlap it, but correctness first: it is a
        // separate pass, and only when the caller asks for it
        end = compress_batch_impl(c, in, sizes.data(), nb, out, out_capacity, csz.data(), hl, tail, bck);
        if (end <= 0) return end ? end : fail(-1, "frame exceeds output capacity %zu", out_capacity);
    }
    memcpy(out, hdr, hl);
    h_wr32((u8*)out + end, 0);                             // end mark, lz4/lz4frame.c:1000-1003
    end += 4;
    if (cck) { h_wr32((u8*)out + end, xxh32_host(in, n)); end += 4; }
    return end;
}

int64_t lz4b200_frame_content_size(const char* in, size_t in_size)
{
    FrameInfo fi; const char* why = "";
    if (!in) return fail(-1, "null argument");
    if (int rc = frame_read_header((const u8*)in, in_size, &fi, &why)) return fail(rc, "%s", why);
    return fi.has_content_size ? (int64_t)fi.content_size : -3;
}

// LZ4F_decompress for a whole frame with independent blocks (lz4/lz4frame.c:1384-1830).  The host walks the block
// size words (one 4-byte read per block); everything that touches block payloads runs on the device: checksum of
// every block, a size pass (the frame does not store decoded sizes), and the chunk decoders.
int64_t lz4b200_frame_decompress(lz4b200_ctx* c, const char* in, size_t in_size, char* out, size_t out_capacity, int flags)
{
    if (!c) return fail(-1, "null ctx");
    if (!in || (!out && out_capacity)) return fail(-1, "null argument");
    CU(cudaSetDevice(c->device));
    FrameInfo fi; const char* why = "";
    const u8* src = (const u8*)in;
    if (int rc = frame_read_header(src, in_size, &fi, &why)) return fail(rc, "%s", why);
    const bool verify = !(flags & LZ4B200_FRAME_SKIP_CHECKSUMS);
    const size_t bx = fi.block_checksum ? 4 : 0;
    // pass 1 (host): block positions
    std::vector<u64> pos; std::vector<u32> words;
    size_t ip = fi.header_bytes;
    for (;;) {
        if (ip + 4 > in_size) return fail(-1, "truncated frame (block header)");
        const u32 w = h_rd32(src + ip);
        ip += 4;
        if (w == 0) break;
        const size_t sz = w & ~F_RAW;
        if (sz > fi.block_max) return fail(-1, "block larger than the frame's block size");
        if (ip + sz + bx > in_size) return fail(-1, "truncated frame (block payload)");
        pos.push_back(ip); words.push_back(w);
        ip += sz + bx;
    }
    u32 stored_cck = 0;
    if (fi.content_checksum) {
        if (ip + 4 > in_size) return fail(-1, "truncated frame (content checksum)");
        stored_cck = h_rd32(src + ip);
        ip += 4;
    }
    const size_t frame_bytes = ip, n = pos.size();
    if (n == 0) {
        if (fi.has_content_size && fi.content_size != 0) return fail(-1, "content size mismatch");
        if (fi.content_checksum && verify && stored_cck != xxh32_host(out, 0)) return fail(-1, "content checksum mismatch");
        return 0;
    }
    if (int rc = check_batch(n, fi.block_max)) return rc;
    maybe_register(c, in, frame_bytes);
    maybe_register(c, out, out_capacity);
    CU(c->h_sizes.ensure((2 * n + lz4b200_ctx::MAX_SLICES + 4) * sizeof(u64)));
    u64* h = (u64*)c->h_sizes.p;
    u32* hw = (u32*)(h + n);
    u64* h_tot = h + 2 * n;
    std::vector<size_t> weight(n);
    for (size_t i = 0; i < n; i++) { h[i] = pos[i]; hw[i] = words[i]; weight[i] = (words[i] & ~F_RAW) + 4 + bx; }
    // decoded bytes are bounded by blocks x block size, and anything beyond the caller's buffer is an error anyway
    const size_t limit = std::min<u64>((u64)n * fi.block_max, out_capacity);
    CU(c->usz.ensure(n * 8)); CU(c->csz.ensure(n * 8)); CU(c->fsz.ensure(n * 8)); CU(c->fhdr.ensure(n * 4));
    CU(c->d_packed.ensure(frame_bytes + 16));
    CU(c->d_plain.ensure(limit + 16));
    if (int rc = ensure_scratch(c, n, 0)) return rc;
    ensure_v7(c, n, n, fi.block_max);
    ensure_coarse(c, n, fi.block_max);
    size_t first[lz4b200_ctx::MAX_SLICES + 1];
    const int K = plan_slices(c, weight.data(), n, frame_bytes, true, first);
    if (int rc = ensure_events(c, K)) return rc;
    u8* sc = (u8*)c->scalars.p;
    u32 *d_bad = (u32*)(sc + SC_BAD), *d_unsup = (u32*)(sc + SC_FUNSUP), *d_ck = (u32*)(sc + SC_FCK);
    u64* running = (u64*)(sc + SC_RUNNING);
    u64* d_pos = (u64*)c->fsz.p;
    CU(cudaMemsetAsync(sc, 0, 32, c->streams[0]));
    CU(cudaMemcpyAsync(d_pos, h, n * 8, cudaMemcpyHostToDevice, c->streams[0]));
    CU(cudaMemcpyAsync(c->fhdr.p, hw, n * 4, cudaMemcpyHostToDevice, c->streams[0]));
    CU(cudaEventRecord(c->ev_sizes, c->streams[0]));
    // slices of blocks flow through H2D -> size pass -> scan (from the previous slice's running total) -> decode;
    // the host learns each slice's output range one event behind and issues the D2H copies on their own stream
    for (int k = 0; k < K; k++) {
        cudaStream_t st = c->streams[k % lz4b200_ctx::NS];
        const size_t f = first[k], m = first[k + 1] - first[k];
        const size_t lo = pos[f] - 4, hi = pos[f + m - 1] + (words[f + m - 1] & ~F_RAW) + bx;
        CU(cudaStreamWaitEvent(st, c->ev_sizes, 0));
        CU(cudaMemcpyAsync((u8*)c->d_packed.p + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, c->own_h2d ? c->h2d_stream : st));
        if (c->own_h2d) { CU(cudaEventRecord(c->ev_in[k], c->h2d_stream)); CU(cudaStreamWaitEvent(st, c->ev_in[k], 0)); }
        u64 *csz_k = (u64*)c->csz.p + f, *usz_k = (u64*)c->usz.p + f;
        frame_block_sizes_kernel<<<(unsigned)((m + FS_WARPS - 1) / FS_WARPS), FS_WARPS * 32, 0, st>>>(
            (const u8*)c->d_packed.p, frame_bytes, d_pos + f, (const u32*)c->fhdr.p + f, csz_k, usz_k, (u32)m,
            (u32)fi.block_max, d_unsup);
        c->launches++;
        if (fi.block_checksum && verify) {
            frame_block_xxh_kernel<<<(unsigned)((m + 63) / 64), 64, 0, st>>>((u8*)c->d_packed.p, d_pos + f, csz_k, (u32)m, 1, d_ck);
            c->launches++;
        }
        const Scratch sk = scratch_at(c, f, 0, k);
        if (k) CU(cudaStreamWaitEvent(st, c->ev_scan[k - 1], 0));
        ScanJob tot{ usz_k, sk.uoff, running + k, k ? running + (k - 1) : nullptr, c->d_mirror + k }, none{ nullptr, nullptr, nullptr, nullptr, nullptr };
        scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>(tot, none, (u64)m);
        CU(cudaEventRecord(c->ev_scan[k], st));
        frame_guard_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(usz_k, csz_k, (u32)m, running + k, (u64)limit);
        c->launches += 2;
        int rc = decompress_slice(c, (const u8*)c->d_packed.p, csz_k, usz_k, m, fi.block_max, (u8*)c->d_plain.p, d_bad,
                                  sk, st, 0, d_pos + f, (u32)fi.block_max);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        CU(cudaEventRecord(c->ev_done[k], st));
    }
    u64 total = 0;
    for (int k = 0; k < K; k++) {
        CU(cudaEventSynchronize(c->ev_done[k]));
        const u64 end = ((volatile u64*)c->h_mirror)[k];
        if (end > out_capacity) {
            CU(cudaDeviceSynchronize());
            return fail(-1, "decoded size exceeds output capacity %zu", out_capacity);
        }
        CU(cudaMemcpyAsync(out + total, (u8*)c->d_plain.p + total, end - total, cudaMemcpyDeviceToHost, c->d2h_stream));
        total = end;
    }
    u32* h_res = (u32*)(h_tot + K);                          // bad, (pad), unsupported, checksum mismatches
    CU(cudaMemcpyAsync(h_res, d_bad, 4, cudaMemcpyDeviceToHost, c->d2h_stream));
    CU(cudaMemcpyAsync(h_res + 2, d_unsup, 8, cudaMemcpyDeviceToHost, c->d2h_stream));
    CU(cudaStreamSynchronize(c->d2h_stream));
    if (h_res[3]) return fail(-1, "%u block checksum mismatch(es)", h_res[3]);
    if (h_res[2]) return fail(-2, "%u compressed block(s) not smaller than their content: not representable on the chunked path", h_res[2]);
    if (h_res[0]) { fail(-1, "%u malformed block(s)", h_res[0]); return -(int64_t)h_res[0]; }
    if (fi.has_content_size && fi.content_size != total) return fail(-1, "content size mismatch");
    if (fi.content_checksum && verify && stored_cck != xxh32_host(out, total)) return fail(-1, "content checksum mismatch");
    return (int64_t)total;
}

}  // extern "C"

