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
#include "staging.cuh"

#include <algorithm>
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
    int kernel_gen = 1;                 // 0 = warp-per-chunk baseline, 1 = tiled CTA kernels
    cudaStream_t stream = nullptr;      // used by the host-buffer entry points
    uint64_t launches = 0;
    // optional per-launch timing of the dominant (codec) kernel: event pairs on the launching stream
    bool timing = false;
    static constexpr int MAX_TIMED = 256;
    cudaEvent_t ev0[MAX_TIMED] = {}, ev1[MAX_TIMED] = {};
    int n_timed = 0;
    // device-resident scratch shared by both directions
    DevBuf usz, csz, uoff, coff;        // u64[n_chunks]
    DevBuf scalars;                     // [0] total (u64), [1] bad chunks (u32), [2] work ticket (u32)
    DevBuf status;                      // u32[n_chunks]
    DevBuf stage;                       // compress: n_chunks slots of slot_stride bytes
    // buffers behind the host entry points
    DevBuf d_plain, d_packed;
    PinBuf h_sizes;                     // u64[2*n_chunks + 2]
};

namespace {

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

// ---- device pipelines -------------------------------------------------------------------

int compress_device_impl(lz4b200_ctx* c, const u8* d_in, const u64* d_usz, size_t n, size_t max_chunk,
                         u8* d_out, u64* d_csz, u64* d_total, cudaStream_t st)
{
    if (n == 0) { CU(cudaMemsetAsync(d_total, 0, sizeof(u64), st)); return 0; }
    if (n > 0x7FFFFFFFull) return fail(-1, "too many chunks");
    if (max_chunk > 0x7E000000ull) return fail(-1, "chunk larger than LZ4_MAX_INPUT_SIZE");
    CU(cudaSetDevice(c->device));
    const size_t stride = slot_stride_for(max_chunk);
    CU(c->uoff.ensure(n * sizeof(u64)));
    CU(c->coff.ensure(n * sizeof(u64)));
    CU(c->stage.ensure(n * stride));
    CU(c->scalars.ensure(64));
    u64* uoff = (u64*)c->uoff.p; u64* coff = (u64*)c->coff.p;

    ScanJob in_scan{ d_usz, uoff, nullptr }, none{ nullptr, nullptr, nullptr };
    scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>(in_scan, none, (u64)n);
    c->launches++;

    time_begin(c, st);
    if (c->kernel_gen >= 1 && v1_compress_supported(max_chunk)) {
        CU(v1_launch_compress(d_in, uoff, d_usz, (u8*)c->stage.p, stride, d_csz, (u32)n,
                              (u32*)c->scalars.p + 4, c->sm_count, st));
        c->launches += 2;
    } else {
        const size_t smem = (size_t)V0_WARPS * V0_TABLE * sizeof(u16);
        static bool attr_set = false;
        if (!attr_set) {
            CU(cudaFuncSetAttribute(compress_v0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        compress_v0_kernel<<<(unsigned)((n + V0_WARPS - 1) / V0_WARPS), V0_WARPS * 32, smem, st>>>(
            d_in, uoff, d_usz, (u8*)c->stage.p, stride, d_csz, (u32)n);
        c->launches++;
    }
    time_end(c, st);
    ScanJob out_scan{ d_csz, coff, d_total };
    scan_sizes_kernel<<<1, SCAN_THREADS, 0, st>>>(out_scan, none, (u64)n);
    c->launches++;
    const unsigned grid = (unsigned)std::min<size_t>(n, (size_t)c->sm_count * 16);
    compact_kernel<<<grid, COMPACT_THREADS, 0, st>>>((const u8*)c->stage.p, stride, d_in, uoff, d_usz,
                                                     d_csz, coff, d_out, (u32)n);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

int decompress_device_impl(lz4b200_ctx* c, const u8* d_in, const u64* d_csz, const u64* d_usz, size_t n,
                           size_t max_chunk, u8* d_out, u32* d_bad, cudaStream_t st)
{
    CU(cudaSetDevice(c->device));
    CU(c->scalars.ensure(64));
    if (!d_bad) d_bad = (u32*)c->scalars.p + 2;
    CU(cudaMemsetAsync(d_bad, 0, sizeof(u32), st));
    if (n == 0) return 0;
    if (n > 0x7FFFFFFFull) return fail(-1, "too many chunks");
    if (max_chunk > 0x7E000000ull) return fail(-1, "chunk larger than LZ4_MAX_INPUT_SIZE");
    CU(c->uoff.ensure(n * sizeof(u64)));
    CU(c->coff.ensure(n * sizeof(u64)));
    CU(c->status.ensure(n * sizeof(u32)));
    u64* uoff = (u64*)c->uoff.p; u64* coff = (u64*)c->coff.p;
    ScanJob a{ d_csz, coff, nullptr }, b{ d_usz, uoff, nullptr };
    scan_sizes_kernel<<<2, SCAN_THREADS, 0, st>>>(a, b, (u64)n);
    c->launches++;
    time_begin(c, st);
    if (c->kernel_gen >= 1 && v1_decompress_supported(max_chunk)) {
        CU(v1_launch_decompress(d_in, coff, d_csz, uoff, d_usz, d_out, (u32*)c->status.p, d_bad, (u32)n,
                                (u32*)c->scalars.p + 4, c->sm_count, st));
        c->launches += 2;
    } else {
        decompress_v0_kernel<<<(unsigned)((n + V0_WARPS - 1) / V0_WARPS), V0_WARPS * 32, 0, st>>>(
            d_in, coff, d_csz, uoff, d_usz, d_out, (u32*)c->status.p, d_bad, (u32)n);
        c->launches++;
    }
    time_end(c, st);
    CU(cudaGetLastError());
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
    if (prop.major < 10) return fail(-2, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    lz4b200_ctx* c = new (std::nothrow) lz4b200_ctx();
    if (!c) return fail(-3, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    const char* gen = getenv("LZ4B200_KERNELS");
    if (gen && !strcmp(gen, "v0")) c->kernel_gen = 0;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(-2, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    if (max_total_bytes && max_chunk_bytes) {
        size_t n = (max_total_bytes + max_chunk_bytes - 1) / max_chunk_bytes;
        // best-effort pre-sizing; growth on demand covers the rest
        c->d_plain.ensure(max_total_bytes);
        c->d_packed.ensure(max_total_bytes + max_total_bytes / 255 + 16 * n);
        c->stage.ensure(n * slot_stride_for(max_chunk_bytes));
        c->usz.ensure(n * 8); c->csz.ensure(n * 8); c->uoff.ensure(n * 8); c->coff.ensure(n * 8);
        c->status.ensure(n * 4);
        c->h_sizes.ensure((2 * n + 2) * 8);
    }
    c->scalars.ensure(64);
    *out = c;
    return 0;
}

void lz4b200_destroy(lz4b200_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->usz.release(); c->csz.release(); c->uoff.release(); c->coff.release();
    c->scalars.release(); c->status.release(); c->stage.release();
    c->d_plain.release(); c->d_packed.release(); c->h_sizes.release();
    for (int i = 0; i < lz4b200_ctx::MAX_TIMED; i++) if (c->ev0[i]) { cudaEventDestroy(c->ev0[i]); cudaEventDestroy(c->ev1[i]); }
    delete c;
}

uint64_t lz4b200_launch_count(const lz4b200_ctx* c) { return c ? c->launches : 0; }

int lz4b200_set_option(lz4b200_ctx* c, const char* key, const char* value)
{
    if (!c || !key || !value) return -1;
    if (!strcmp(key, "kernels")) {
        if (!strcmp(value, "v0")) { c->kernel_gen = 0; return 0; }
        if (!strcmp(value, "v1")) { c->kernel_gen = 1; return 0; }
    }
    if (!strcmp(key, "timing")) { c->timing = value[0] == '1'; c->n_timed = 0; return 0; }
    return fail(-1, "unknown option %s=%s", key, value);
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

int lz4b200_compress_device(lz4b200_ctx* c, const void* d_in, const uint64_t* d_chunk_sizes, size_t n_chunks,
                            size_t max_chunk_bytes, void* d_out, uint64_t* d_compr_sizes,
                            uint64_t* d_total_out, void* stream)
{
    if (!c) return fail(-1, "null ctx");
    return compress_device_impl(c, (const u8*)d_in, (const u64*)d_chunk_sizes, n_chunks, max_chunk_bytes,
                                (u8*)d_out, (u64*)d_compr_sizes, (u64*)d_total_out, (cudaStream_t)stream);
}

int lz4b200_decompress_device(lz4b200_ctx* c, const void* d_in, const uint64_t* d_compr_sizes,
                              const uint64_t* d_chunk_sizes, size_t n_chunks, size_t max_chunk_bytes,
                              void* d_out, uint32_t* d_bad_chunks, void* stream)
{
    if (!c) return fail(-1, "null ctx");
    return decompress_device_impl(c, (const u8*)d_in, (const u64*)d_compr_sizes, (const u64*)d_chunk_sizes,
                                  n_chunks, max_chunk_bytes, (u8*)d_out, d_bad_chunks, (cudaStream_t)stream);
}

int64_t lz4b200_compress_batch(lz4b200_ctx* c, const char* in, const size_t* chunk_sizes, size_t n,
                               char* out, size_t out_capacity, size_t* compr_sizes)
{
    if (!c) return fail(-1, "null ctx");
    if (n == 0) return 0;
    if (!in || !chunk_sizes || !out || !compr_sizes) return fail(-1, "null argument");
    CU(cudaSetDevice(c->device));
    size_t total = 0, max_chunk = 0;
    CU(c->h_sizes.ensure((2 * n + 2) * sizeof(u64)));
    u64* h = (u64*)c->h_sizes.p;
    for (size_t i = 0; i < n; i++) { h[i] = chunk_sizes[i]; total += chunk_sizes[i]; max_chunk = std::max(max_chunk, chunk_sizes[i]); }
    if (max_chunk > 0x7E000000ull) return fail(-1, "chunk larger than LZ4_MAX_INPUT_SIZE");
    CU(c->usz.ensure(n * 8)); CU(c->csz.ensure(n * 8));
    CU(c->d_plain.ensure(total + 16));
    CU(c->d_packed.ensure(total + total / 255 + 16 * n + 16));
    CU(c->scalars.ensure(64));
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->usz.p, h, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_plain.p, in, total, cudaMemcpyHostToDevice, st));
    int rc = compress_device_impl(c, (const u8*)c->d_plain.p, (const u64*)c->usz.p, n, max_chunk,
                                  (u8*)c->d_packed.p, (u64*)c->csz.p, (u64*)c->scalars.p, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h + n, c->csz.p, n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h + 2 * n, c->scalars.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const u64 ctotal = h[2 * n];
    if (ctotal > out_capacity) { fail(0, "compressed size %llu exceeds output capacity %zu", ctotal, out_capacity); return 0; }
    CU(cudaMemcpyAsync(out, c->d_packed.p, ctotal, cudaMemcpyDeviceToHost, st));
    for (size_t i = 0; i < n; i++) compr_sizes[i] = (size_t)h[n + i];
    CU(cudaStreamSynchronize(st));
    return (int64_t)ctotal;
}

int64_t lz4b200_decompress_batch(lz4b200_ctx* c, const char* in, const size_t* compr_sizes,
                                 const size_t* chunk_sizes, size_t n, char* out)
{
    if (!c) return fail(-1, "null ctx");
    if (n == 0) return 0;
    if (!in || !chunk_sizes || !out || !compr_sizes) return fail(-1, "null argument");
    CU(cudaSetDevice(c->device));
    size_t total = 0, ctotal = 0, max_chunk = 0;
    CU(c->h_sizes.ensure((2 * n + 2) * sizeof(u64)));
    u64* h = (u64*)c->h_sizes.p;
    for (size_t i = 0; i < n; i++) {
        h[i] = chunk_sizes[i]; h[n + i] = compr_sizes[i];
        total += chunk_sizes[i]; ctotal += compr_sizes[i];
        max_chunk = std::max(max_chunk, chunk_sizes[i]);
    }
    if (max_chunk > 0x7E000000ull) return fail(-1, "chunk larger than LZ4_MAX_INPUT_SIZE");
    CU(c->usz.ensure(n * 8)); CU(c->csz.ensure(n * 8));
    CU(c->d_plain.ensure(total + 16));
    CU(c->d_packed.ensure(ctotal + 16));
    CU(c->scalars.ensure(64));
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->usz.p, h, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->csz.p, h + n, n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_packed.p, in, ctotal, cudaMemcpyHostToDevice, st));
    u32* d_bad = (u32*)c->scalars.p + 2;
    int rc = decompress_device_impl(c, (const u8*)c->d_packed.p, (const u64*)c->csz.p, (const u64*)c->usz.p, n,
                                    max_chunk, (u8*)c->d_plain.p, d_bad, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, c->d_plain.p, total, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h + 2 * n, d_bad, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const u32 bad = *(const u32*)(h + 2 * n);
    if (bad) { fail(-1, "%u malformed chunk(s)", bad); return -(int64_t)bad; }
    return (int64_t)total;
}

}  // extern "C"
