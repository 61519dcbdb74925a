// rt.cuh -- the one place that knows whether the sources are compiled by nvcc for sm_100a (the
// product) or by g++ against tests/hostsim/cusim.h (RSI_SIM: the CPU emulator the test-suite uses
// because the development container has no GPU).  The product library never defines RSI_SIM.
#pragma once
#include <stdint.h>

#if defined(RSI_SIM)
#include "cusim.h"
#define RSI_LAUNCH(kern, grid, block, smem, stream, ...) \
  do { (void)(stream); cusim::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); }); } while (0)
#define RSI_DYN_SMEM(name) unsigned char* name = cusim::S().dyn()
#define RSI_LAUNCH_CLUSTER(kern, grid, block, cluster, smem, stream, ...) \
  do { (void)(stream); cusim::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); }, (cluster)); } while (0)
#else
#include <cuda_runtime.h>
#define RSI_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<dim3(grid), dim3(block), (size_t)(smem), (stream)>>>(__VA_ARGS__)
#define RSI_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
// launch with a thread-block cluster of `cluster` CTAs along x (grid.x must be a multiple of it)
#define RSI_LAUNCH_CLUSTER(kern, grid, block, cluster, smem, strm_, ...)                                        \
  do {                                                                                                          \
    cudaLaunchConfig_t cfg_ = {};                                                                               \
    cfg_.gridDim = dim3(grid); cfg_.blockDim = dim3(block); cfg_.dynamicSmemBytes = (size_t)(smem); cfg_.stream = (strm_); \
    cudaLaunchAttribute at_[1];                                                                                 \
    at_[0].id = cudaLaunchAttributeClusterDimension;                                                            \
    at_[0].val.clusterDim.x = (cluster); at_[0].val.clusterDim.y = 1; at_[0].val.clusterDim.z = 1;              \
    cfg_.attrs = at_; cfg_.numAttrs = 1;                                                                        \
    cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__);                                                               \
  } while (0)
#endif

namespace rsigpu {
typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ i64 lmin(i64 a, i64 b) { return a < b ? a : b; }
__device__ __forceinline__ i64 lmax(i64 a, i64 b) { return a > b ? a : b; }
__device__ __forceinline__ int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// order-preserving float <-> unsigned maps for atomicMin / atomicMax on floats
__device__ __forceinline__ u32 f2ord(float f) { u32 u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(u32 u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
}  // namespace rsigpu
