// kernels_v1.cuh -- generation-1 tiled CTA kernels (placeholder until the first v0 numbers are in).
#pragma once
#include "common.cuh"

namespace lz4b200 {

inline bool v1_compress_supported(size_t) { return false; }
inline bool v1_decompress_supported(size_t) { return false; }

inline cudaError_t v1_launch_compress(const u8*, const u64*, const u64*, u8*, u64, u64*, u32, u32*, int, cudaStream_t)
{ return cudaErrorNotSupported; }
inline cudaError_t v1_launch_decompress(const u8*, const u64*, const u64*, const u64*, const u64*, u8*, u32*, u32*, u32,
                                        u32*, int, cudaStream_t)
{ return cudaErrorNotSupported; }

}  // namespace lz4b200
