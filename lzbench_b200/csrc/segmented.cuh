// segmented.cuh -- chunks larger than 64 KiB on the tiled compressor.
//
// The tiled compressor (kernels_c2.cuh) keeps one 64 KiB window in shared memory, so a larger chunk is cut into
// SEGMENTS of 48 KiB of new data that are compressed as independent jobs -- each job also loads the 16 KiB before
// its segment as history (hashed and matchable, not encoded; most LZ4 matches are that close), which keeps the
// size within the tolerance of the reference's sliding 64 KiB window -- and then STITCHED into the single LZ4 block the chunk must be (lzbench's compbuf has no per-segment
// framing).  A segment's stand-alone block is   seq_1 ... seq_k  last_literals(T);   inside a chunk the T tail
// literals of segment s are not a sequence of their own: they become the head of seq_1's literal run in the next
// segment that has a match.  Literals are contiguous in the INPUT, so the merged run is copied from the input
// and only seq_1's token / length bytes are re-encoded:
//     piece(s) = token'(carry + L1, m1)  len-bytes(carry + L1)  input[seg_start - carry, seg_start + L1)  body(s)
//     body(s)  = the segment's block after seq_1's literals, up to its last_literals record
//     carry(s+1) = T(s) if the segment had a match, else carry(s) + seg_len(s)
// and the chunk ends with one last_literals record for the final carry (lz4/lz4.c:1204-1231).
#pragma once
#include "common.cuh"
#include "kernels_c1.cuh"

namespace lz4b200 {

constexpr u32 SEG_BYTES = 65536;                 // bytes a job holds in shared memory
constexpr u32 SEG_HIST = 16384;                  // of which history: the previous 16 KiB of the chunk, match source only
constexpr u32 SEG_TAIL_MIN = 16;                  // fewest new bytes the last segment of a chunk may hold
constexpr u32 SEG_NEW = SEG_BYTES - SEG_HIST;    // and new data, which the job encodes

struct SegArrays {
    u64* jobs_per_chunk;   // [n_chunks]   number of segments of each chunk
    u64* job_base;         // [n_chunks]   first job of each chunk (exclusive scan)
    u64* n_jobs;           // [1]          total jobs (scan total)
    u64* job_in_off;       // [jobs]       job start (history included), relative to the slice's input
    u64* job_len;          // [jobs]       job bytes (history + new data)
    u32* job_hist;         // [jobs]       history bytes at the start of the job; bit 31 = not the chunk's last segment
    u64* job_csz;          // [jobs]       size of the segment's stand-alone block (written by the compressor)
    u32* job_tail_pos;     // [jobs]       offset of the last_literals record inside that block
    u32* job_tail_len;     // [jobs]       T: number of tail literals
    u32* job_rel;          // [jobs]       where the segment's piece starts inside the chunk's final block
    u32* job_carry;        // [jobs]       literals carried into the segment
    u32* job_chunk;        // [jobs]
};

__global__ void seg_count_kernel(const u64* __restrict__ usz, u32 n_chunks, u64* __restrict__ jobs_per_chunk)
{
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const u64 n = usz[c];
    jobs_per_chunk[c] = n ? (n + SEG_NEW - 1) / SEG_NEW : 1;
}

__global__ void seg_fill_kernel(const u64* __restrict__ usz, const u64* __restrict__ uoff, SegArrays a, u32 n_chunks)
{
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const u64 n = usz[c], base = a.job_base[c], cnt = a.jobs_per_chunk[c];
    // An inner segment may match right up to its last byte, so the chunk's end-of-block rules (last match starts >= 12 and
    // ends >= 5 bytes before the end, lz4/lz4.c:883-884) rest on the LAST segment alone: it must not be shorter than 16 new
    // bytes.  If the cut would leave fewer, the last boundary moves 16 bytes down (found by scripts/fuzz_gpu.py: a chunk of
    // 49 153 bytes came out with one literal behind its last match).
    const u64 shift = (cnt > 1 && n - (cnt - 1) * SEG_NEW < SEG_TAIL_MIN) ? SEG_TAIL_MIN : 0;
    for (u64 s = 0; s < cnt; s++) {
        const u64 off = s * SEG_NEW - (s + 1 == cnt ? shift : 0);
        const u32 hist = s ? SEG_HIST : 0;                     // SEG_NEW >= SEG_HIST: the history always exists
        const u64 fresh = s + 1 == cnt ? n - off : SEG_NEW - (s + 2 == cnt ? shift : 0);
        a.job_in_off[base + s] = uoff[c] + off - hist;
        a.job_len[base + s] = hist + fresh;
        a.job_hist[base + s] = hist | (s + 1 < cnt ? 0x80000000u : 0u);   // bit 31: not the last segment of its chunk
        a.job_chunk[base + s] = c;
    }
}

// literal length and header size (token + length bytes) of the first sequence of a block
__device__ __forceinline__ void seg_first_seq(const u8* blk, u32& L1, u32& H1)
{
    u32 len = blk[0] >> 4, h = 1;
    if (len == RUN_MASK) { u32 b; do { b = blk[h++]; len += b; } while (b == 255); }
    L1 = len; H1 = h;
}

// One thread per chunk: walk its segments, fix every piece's place and the chunk's compressed size.
__global__ void seg_stitch_size_kernel(SegArrays a, const u8* __restrict__ stage, u64 slot_stride,
                                       u64* __restrict__ csz, u32 n_chunks)
{
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const u64 base = a.job_base[c], cnt = a.jobs_per_chunk[c];
    u32 carry = 0, out = 0;
    for (u64 s = 0; s < cnt; s++) {
        const u64 j = base + s;
        a.job_rel[j] = out; a.job_carry[j] = carry;
        const u32 tail = a.job_tail_pos[j];
        if (tail) {
            u32 L1, H1;
            seg_first_seq(stage + j * slot_stride, L1, H1);
            const u32 lit = carry + L1;
            out += 1 + c1_len_ext_bytes(lit) + lit + (tail - H1 - L1);
            carry = a.job_tail_len[j];
        } else {
            carry += (u32)a.job_len[j] - (a.job_hist[j] & 0x7FFFFFFFu);
        }
    }
    csz[c] = (u64)out + 1 + c1_len_ext_bytes(carry) + carry;
}

constexpr int SEG_COPY_THREADS = 256;

// One CTA per segment (grid-stride): write its piece; the chunk's last segment also writes the final
// last_literals record.  A chunk whose block came out exactly chunk-sized is stored raw
// (_lzbench/lzbench.cpp:284-289): its last segment copies the input instead and the others write nothing.
__global__ void __launch_bounds__(SEG_COPY_THREADS)
seg_stitch_copy_kernel(SegArrays a, const u8* __restrict__ stage, u64 slot_stride, const u8* __restrict__ in,
                       const u64* __restrict__ uoff, const u64* __restrict__ usz, const u64* __restrict__ csz,
                       const u64* __restrict__ coff, u8* __restrict__ out, u32 max_jobs)
{
    const u32 n_jobs = (u32)*a.n_jobs;
    const u32 tid = threadIdx.x;
    for (u32 j = blockIdx.x; j < n_jobs && j < max_jobs; j += gridDim.x) {
        const u32 c = a.job_chunk[j];
        const bool last_seg = j + 1 == a.job_base[c] + a.jobs_per_chunk[c];
        u8* dst = out + coff[c];
        if (csz[c] == usz[c]) {                                    // raw-stored chunk
            if (last_seg) coop_copy(dst, in + uoff[c], usz[c], tid, SEG_COPY_THREADS);
            continue;
        }
        const u8* blk = stage + (u64)j * slot_stride;
        const u8* seg_in = in + a.job_in_off[j] + (a.job_hist[j] & 0x7FFFFFFFu);   // first byte of the segment's new data
        u32 o = a.job_rel[j];
        u32 carry = a.job_carry[j];
        const u32 tail = a.job_tail_pos[j];
        if (tail) {
            u32 L1, H1;
            seg_first_seq(blk, L1, H1);
            const u32 lit = carry + L1;
            const u32 hdr = 1 + c1_len_ext_bytes(lit);
            if (tid == 0) {
                dst[o] = (u8)(((lit < RUN_MASK ? lit : RUN_MASK) << 4) | (blk[0] & ML_MASK));
                if (lit >= RUN_MASK) c1_put_ext(dst + o + 1, lit - RUN_MASK);
            }
            coop_copy(dst + o + hdr, seg_in - carry, lit, tid, SEG_COPY_THREADS);
            coop_copy(dst + o + hdr + lit, blk + H1 + L1, tail - H1 - L1, tid, SEG_COPY_THREADS);
            o += hdr + lit + (tail - H1 - L1);
            carry = a.job_tail_len[j];
        } else {
            carry += (u32)a.job_len[j] - (a.job_hist[j] & 0x7FFFFFFFu);
        }
        if (last_seg) {                                            // final literal-only sequence of the chunk
            const u32 hdr = 1 + c1_len_ext_bytes(carry);
            if (tid == 0) {
                dst[o] = (u8)((carry < RUN_MASK ? carry : RUN_MASK) << 4);
                if (carry >= RUN_MASK) c1_put_ext(dst + o + 1, carry - RUN_MASK);
            }
            coop_copy(dst + o + hdr, in + uoff[c] + usz[c] - carry, carry, tid, SEG_COPY_THREADS);
        }
    }
}

}  // namespace lz4b200
