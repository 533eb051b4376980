// kernels_c1.cuh -- generation-1 compressor (placeholder while the host pipeline is verified).
#pragma once
#include "common.cuh"
namespace lz4b200 {
inline bool c1_compress_supported(size_t) { return false; }
inline cudaError_t c1_launch_compress(const u8*, const u64*, const u64*, u8*, u64, u64*, u32, u32*, int, size_t, cudaStream_t)
{ return cudaErrorNotSupported; }
}
