// kernels_coarse.cuh -- decode of LARGE chunks that consist of a handful of long copies (incompressible data: one literal
// run; run-length / short-period data: one match that repeats the bytes before it; raw-stored chunks).
//
// Every decoder of this library gives a chunk to one team of lanes, which is what a chunk of thousands of short
// sequences needs.  A 4 MiB chunk of random bytes is ONE sequence, and 1 GiB of them is 256 chunks: 256 teams copy at
// 107 GB/s on a machine that copies at 1.8 TB/s when the same data comes in 64 KiB chunks (profiles/r02_chunk_sweep_1gib.jsonl,
// BASELINE configs[4]).  For chunks above 64 KiB that can only hold few sequences (the same test kernels_v7.cuh routes by):
//   1. coarse_scan_kernel, one warp per chunk, walks the token chain (lz4/lz4.c:1797-2151; runs of 255 length bytes 128 at
//      a time) and turns it into at most C_MAXD copy descriptors.  It ACCEPTS only: every rule of the reference decoder
//      must hold with room to spare, anything else -- too many sequences, any doubt about a limit -- leaves the chunk on
//      the list of the team decoder, which names the violated rule as before.
//      Every descriptor gets a LEVEL: literal runs 0 (they read the stream), a match 1 + the highest level among the earlier
//      descriptors whose output its source touches.
//   2. coarse_copy_kernel, one persistent launch, one round per level with a grid-wide barrier between rounds: in a round every
//      32 KiB piece of every descriptor of that level is copied by some CTA.  A literal run is a plain copy; a match of
//      length L at offset d is out[p + i] = out[p - d + i mod d] -- all of it reads the d bytes in front of it, which earlier
//      rounds have completed.  Incompressible data has a few dozen accidental 4-byte matches per MiB, all of level 1: two
//      rounds.  Run-length data from this library's segmented compressor is a chain of one match per 48 KiB segment: a round
//      per match (12 MiB per round on 256 chunks of 4 MiB).
#pragma once
#include "common.cuh"

namespace lz4b200 {

constexpr u32 C_MAXD = 256;                      // descriptors per chunk (128 sequences)
constexpr u32 C_PIECE = 32768;                   // bytes of a descriptor one CTA copies at a time
constexpr u32 C_COPY_THREADS = 256;
constexpr u32 C_SCAN_WARPS = 4;

struct CoarseParams {
    const u8* in; const u64* in_off; const u64* csz; const u64* out_off; const u64* usz; u8* out;
    u32* status; u32 n_chunks;
    u32 cap_all;                                   // see V5Params::cap_all (frames: the end-of-block rules are checked against the frame's block size)
    uint4* desc;                                   // [n_chunks][C_MAXD]: x position in the chunk, y stream position (literals) / position of the source period (match), z length, w bit 0: match, bits 8..: level
    uint2* desc2;                                  // [n_chunks][C_MAXD], matches: x period d, y phase: out[x + i] = out[y_source + (phase + i) mod d]
    u32* ndesc;                                    // descriptors of the chunk (0: not taken)
    u32* list; u32* list_count;                    // chunks left to the team decoder
    u32* max_level;                                // highest level of the launch; behind it the grid barrier's counter
};

// value of a length that continues at stream position ip (the nibble was 15): sum of the bytes up to and including the first
// one below 255 (lz4/lz4.c:1707-1729).  128 bytes per warp step.  Returns false if the stream ends first.
__device__ __forceinline__ bool coarse_read_length(const u8* src, u32 clen, u32& ip, u64& len, u32 lane)
{
    for (;;) {
        u32 b[4]; bool stop = false; u32 first = 4;
        #pragma unroll
        for (u32 j = 0; j < 4; j++) {
            const u32 x = ip + 4 * lane + j;
            b[j] = x < clen ? (u32)__ldg(src + x) : 0x100u;              // 0x100: beyond the stream
            if (!stop && b[j] != 255u) { stop = true; first = j; }
        }
        const unsigned m = __ballot_sync(FULL, stop);
        if (!m) { len += 128u * 255u; ip += 128; continue; }
        const u32 fl = (u32)__ffs(m) - 1u;
        const u32 fj = __shfl_sync(FULL, first, fl);
        u32 v = 0;
        #pragma unroll
        for (u32 j = 0; j < 4; j++) if (j == fj) v = b[j];
        v = __shfl_sync(FULL, v, fl);
        if (v > 255u) return false;
        len += (u64)(4 * fl + fj) * 255u + v;
        ip += 4 * fl + fj + 1;
        return true;
    }
}

__global__ void __launch_bounds__(C_SCAN_WARPS * 32)
coarse_scan_kernel(const __grid_constant__ CoarseParams P)
{
    __shared__ u32 s_pos[C_SCAN_WARPS][C_MAXD], s_len[C_SCAN_WARPS][C_MAXD], s_lvl[C_SCAN_WARPS][C_MAXD];   // matches of the chunk so far: output range, level,
    __shared__ u32 s_src[C_SCAN_WARPS][C_MAXD], s_dph[C_SCAN_WARPS][C_MAXD];                               // ... source period (position; d | phase << 16)
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 c = blockIdx.x * C_SCAN_WARPS + w;
    if (c >= P.n_chunks) return;
    const u8* src = P.in + P.in_off[c];
    const u64 clen64 = P.csz[c], n64 = P.usz[c];
    uint4* const d = P.desc + (size_t)c * C_MAXD;
    uint2* const d2 = P.desc2 + (size_t)c * C_MAXD;
    u32 nd = 0, nm = 0, top = 0;
    bool take = false;
    if (clen64 == n64 && n64 != 0 && n64 < 0xFFFFFFFFull) {                 // raw chunk (_lzbench/lzbench.cpp:311-315): a plain copy
        if (lane == 0) d[0] = make_uint4(0u, 0u, (u32)n64, 0u);
        nd = 1; take = true;
    } else if (n64 >= 65536 && n64 < 0x7F000000ull && clen64 >= 16 && clen64 < 0x7F000000ull &&
               (clen64 * 85 < n64 || clen64 + (n64 >> 5) >= n64)) {          // few sequences: fewer than n / 256, or hardly compressed
        const u32 clen = (u32)clen64, n = (u32)n64;
        const u32 capv = P.cap_all > n ? P.cap_all : n;
        u32 ip = 0, op = 0;
        take = true;
        while (take) {
            if (nd + 2 > C_MAXD || ip >= clen) { take = false; break; }      // more sequences than this path is for
            const u32 tk = __ldg(src + ip); ip++;
            u64 len = tk >> 4;
            if (len == RUN_MASK) {
                if (!coarse_read_length(src, clen, ip, len, lane) || ip + RUN_MASK >= clen) { take = false; break; }
            }
            if (len > n) { take = false; break; }
            const u32 L = (u32)len, cpy = op + L;
            if (cpy > n || (u64)ip + L > clen) { take = false; break; }
            const bool last = cpy + MFLIMIT > capv || (u64)ip + L + (2 + 1 + LASTLITERALS) > clen;
            if (L) { if (lane == 0) d[nd] = make_uint4(op, ip, L, 0u); nd++; }
            if (last) {
                if (ip + L != clen || cpy != n) take = false;                // (the team decoder names the rule)
                break;
            }
            ip += L;
            const u32 off = (u32)__ldg(src + ip) | ((u32)__ldg(src + ip + 1) << 8);
            ip += 2;
            u64 m = tk & ML_MASK;
            if (m == ML_MASK) {
                if (!coarse_read_length(src, clen, ip, m, lane) || ip + LASTLITERALS - 1 >= clen) { take = false; break; }
            }
            m += MINMATCH;
            if (m > n || cpy + (u32)m + LASTLITERALS > capv || cpy + (u32)m > n || off == 0 || off > cpy) { take = false; break; }
            // level: one above the highest earlier match whose output the source [cpy - off, cpy - off + min(off, m)) touches.
            // FORWARDING: if the whole source lies inside the output of ONE earlier match k whose period divides this offset, the
            // bytes are k's period again -- out[cpy + i] = out[src_k + (phase + i) mod d_k] -- so this match reads what k reads and
            // runs in k's round (run-length data compressed in segments is a chain of such matches: one round instead of one each)
            const u32 s0 = cpy - off, s1 = s0 + (off < (u32)m ? off : (u32)m);
            u32 lvl = 0, fk = 0xFFFFFFFFu;
            for (u32 k = lane; k < nm; k += 32) {
                const u32 pk = s_pos[w][k], ek = pk + s_len[w][k];
                if (pk < s1 && ek > s0 && s_lvl[w][k] > lvl) lvl = s_lvl[w][k];
                if (pk <= s0 && s1 <= ek && off % (s_dph[w][k] & 0xFFFFu) == 0) fk = k;
            }
            #pragma unroll
            for (u32 dd = 16; dd; dd >>= 1) { const u32 y = __shfl_xor_sync(FULL, lvl, dd); if (y > lvl) lvl = y; }
            lvl += 1;
            const unsigned fm = __ballot_sync(FULL, fk != 0xFFFFFFFFu);
            u32 srcp = s0, per = off, ph = 0;
            if (fm) {
                fk = __shfl_sync(FULL, fk, __ffs(fm) - 1);
                per = s_dph[w][fk] & 0xFFFFu;
                ph = ((s_dph[w][fk] >> 16) + (s0 - s_pos[w][fk])) % per;
                srcp = s_src[w][fk]; lvl = s_lvl[w][fk];
            }
            if (lane == 0) {
                d[nd] = make_uint4(cpy, srcp, (u32)m, 1u | (lvl << 8));
                d2[nd] = make_uint2(per, ph);
                s_pos[w][nm] = cpy; s_len[w][nm] = (u32)m; s_lvl[w][nm] = lvl; s_src[w][nm] = srcp; s_dph[w][nm] = per | (ph << 16);
            }
            __syncwarp();
            nd++; nm++;
            if (lvl > top) top = lvl;
            op = cpy + (u32)m;
        }
    }
    if (lane == 0) {
        P.ndesc[c] = take ? nd : 0u;
        if (take) { P.status[c] = ST_OK; atomicMax(P.max_level, top); }
        else P.list[atomicAdd(P.list_count, 1u)] = c;
    }
}

// all CTAs of the launch are resident (the grid is sized by occupancy and the launch is cooperative): a counter that only grows is the barrier
__device__ __forceinline__ void coarse_grid_barrier(u32* counter, u32& round)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const u32 target = (round + 1) * gridDim.x;
        while (*reinterpret_cast<volatile u32*>(counter) < target) __nanosleep(100);
        __threadfence();
    }
    round++;
    __syncthreads();
}

// one round per level; in a round, the pieces of the descriptors of that level.  CTA b works on the chunks c = b mod n_chunks
// (+ multiples of the grid size); the CTAs that share a chunk share its pieces.
__global__ void __launch_bounds__(C_COPY_THREADS)
coarse_copy_kernel(const __grid_constant__ CoarseParams P)
{
    __shared__ u32 s_mask[C_MAXD / 32];                 // descriptors of the current level (in descriptor order: the CTAs that share a chunk number its pieces alike)
    const u32 G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const u32 levels = *reinterpret_cast<volatile u32*>(P.max_level) + 1;
    u32 round = 0;
    for (u32 lv = 0; lv < levels; lv++) {
        for (u32 c = b % P.n_chunks; c < P.n_chunks; c += G) {
            const u32 nd = P.ndesc[c];
            if (nd) {
                // CTAs of this chunk: b, b + n, b + 2n ... < G when the grid is larger than the chunk count; else this one alone
                const u32 share = G > P.n_chunks ? (G - 1 - (b % P.n_chunks)) / P.n_chunks + 1 : 1;
                const u32 mine = G > P.n_chunks ? b / P.n_chunks : 0;
                const uint4* const dsc = P.desc + (size_t)c * C_MAXD;
                // the descriptors of this level, in order (one thread each looks at its descriptor)
                {
                    const unsigned bal = __ballot_sync(FULL, tid < nd && (dsc[tid < nd ? tid : 0].w >> 8) == lv);
                    if ((tid & 31) == 0) s_mask[tid >> 5] = bal;
                }
                __syncthreads();
                u32 piece = 0;                                             // running piece number over the level's descriptors of the chunk
                for (u32 k = 0; k < nd; k++) {
                    if (!((s_mask[k >> 5] >> (k & 31)) & 1u)) continue;
                    const uint4 q = dsc[k];
                    u8* const dst = P.out + P.out_off[c] + q.x;
                    for (u64 a = 0; a < q.z; a += C_PIECE, piece++) {
                        if (piece % share != mine) continue;
                        const u32 len = q.z - a < C_PIECE ? (u32)(q.z - a) : C_PIECE;
                        if ((q.w & 1u) == 0) {
                            coop_copy(dst + a, P.in + P.in_off[c] + q.y + a, len, tid, C_COPY_THREADS);
                        } else {
                            // out[p + i] = out[p - d + i mod d]: 16-byte stores once the target is aligned, sources byte by byte from
                            // the period in front of the match (L1 / L2 hits)
                            const uint2 q2 = P.desc2[(size_t)c * C_MAXD + k];
                            const u32 dd = q2.x;
                            const u8* const per = P.out + P.out_off[c] + q.y;
                            const u64 a0 = a + q2.y;                                 // phase of this piece's first byte
                            u8* const t = dst + a;
                            u32 head = (u32)((16 - (reinterpret_cast<uintptr_t>(t) & 15)) & 15);
                            if (head > len) head = len;
                            if (tid < head) t[tid] = per[(a0 + tid) % dd];
                            const u32 nvec = (len - head) >> 4;
                            for (u32 j = tid; j < nvec; j += C_COPY_THREADS) {
                                u32 s = (u32)((a0 + head + (u64)j * 16) % dd);
                                u32 wv[4];
                                #pragma unroll
                                for (u32 kk = 0; kk < 4; kk++) {
                                    u32 x = 0;
                                    #pragma unroll
                                    for (u32 bb = 0; bb < 4; bb++) { x |= (u32)per[s] << (8 * bb); s++; if (s == dd) s = 0; }
                                    wv[kk] = x;
                                }
                                *reinterpret_cast<uint4*>(t + head + (size_t)j * 16) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                            }
                            const u32 done = head + nvec * 16;
                            if (tid < len - done) t[done + tid] = per[(a0 + done + tid) % dd];
                        }
                    }
                }
                __syncthreads();
            }
        }
        if (lv + 1 < levels) coarse_grid_barrier(P.max_level + 1, round);
    }
}

inline cudaError_t coarse_launch(const CoarseParams& P, int sm_count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(P.list_count, 0, sizeof(u32), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.max_level, 0, 2 * sizeof(u32), st);
    if (e != cudaSuccess) return e;
    coarse_scan_kernel<<<(P.n_chunks + C_SCAN_WARPS - 1) / C_SCAN_WARPS, C_SCAN_WARPS * 32, 0, st>>>(P);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coarse_copy_kernel, C_COPY_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;                          // (the barrier needs every CTA resident: room is left for whatever else runs)
    // a cooperative launch: the runtime starts the kernel only when all its CTAs can be resident together (the barrier's premise)
    CoarseParams Pc = P;
    void* args[] = { (void*)&Pc };
    e = cudaLaunchCooperativeKernel((const void*)coarse_copy_kernel, dim3((unsigned)(sm_count * per_sm)), dim3(C_COPY_THREADS), args, 0, st);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace lz4b200
