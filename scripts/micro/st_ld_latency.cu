// Microbenchmark for DESIGN.md "what next": how long does a byte load take when the same lane stored that byte a few
// instructions earlier (the common case of an LZ4 match with a short offset), compared with a byte that was written
// long ago (L2 hit) and one that is not in L2?   nvcc -arch=sm_100a -O3 -o st_ld_latency st_ld_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint8_t* buf, size_t far_stride, long long* out, int mode, int iters)
{
    const int lane = threadIdx.x & 31;
    uint8_t* p = buf + (size_t)blockIdx.x * (1 << 20) + lane;
    long long acc = 0; unsigned v = lane;
    for (int i = 0; i < iters; i++) {
        uint8_t* q = p + (size_t)i * 64;
        if (mode == 0) q[0] = (uint8_t)v;                       // store, then load the same byte
        const uint8_t* src = mode == 0 ? q : (mode == 1 ? p + (size_t)(i & 63) * 64 : buf + ((size_t)(i * 7919 + blockIdx.x * 104729) % far_stride) * 4096 + lane);
        const long long t0 = clock64();
        v += *(volatile const uint8_t*)src;                      // dependent use
        if (v == 0x12345678u) q[1] = 1;
        const long long t1 = clock64();
        acc += t1 - t0;
    }
    if (lane == 0) out[blockIdx.x] = acc / iters;
    if (v == 0xdeadbeef) out[0] = v;
}
int main()
{
    const size_t bytes = (size_t)4 << 30;
    uint8_t* buf; long long* out; cudaMalloc(&buf, bytes); cudaMalloc(&out, 1024 * 8); cudaMemset(buf, 1, bytes);
    const char* names[3] = { "load right after own store", "load of a byte stored long ago (L1/L2 hit)", "load of a cold byte (mostly HBM)" };
    for (int blocks : { 1, 148 * 7 }) for (int mode = 0; mode < 3; mode++) {
        k<<<blocks, 128>>>(buf, bytes / 4096 - 1, out, mode, 2000); cudaDeviceSynchronize();
        long long h[1024]; cudaMemcpy(h, out, sizeof(long long) * (blocks < 1024 ? blocks : 1024), cudaMemcpyDeviceToHost);
        long long s = 0; int n = blocks < 1024 ? blocks : 1024; for (int i = 0; i < n; i++) s += h[i];
        printf("%4d CTAs x 4 warps: %-46s %5lld cycles\n", blocks, names[mode], s / n);
    }
    return 0;
}
