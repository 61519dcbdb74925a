// rt.cuh -- the one place that knows whether the sources are compiled by nvcc for sm_100a (the
// product) or by g++ against tests/hostsim/cusim.h (RSI_SIM: the CPU emulator the test-suite uses
// because the development container has no GPU).  The product library never defines RSI_SIM.
#pragma once
#include <stdint.h>

#if defined(RSI_SIM)
#include "cusim.h"
#define RSI_LAUNCH(kern, grid, block, smem, stream, ...) \
  do { (void)(stream); cusim::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); }); } while (0)
#define RSI_DYN_SMEM(name) unsigned char* name = cusim::S().dyn_smem
#else
#include <cuda_runtime.h>
#define RSI_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<dim3(grid), dim3(block), (size_t)(smem), (stream)>>>(__VA_ARGS__)
#define RSI_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
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

// 16-byte asynchronous global -> shared copy (LDGSTS): many copies in flight per thread without staging registers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
#if defined(RSI_SIM)
  *reinterpret_cast<uint4*>(smem_dst) = *reinterpret_cast<const uint4*>(gsrc);
#else
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#if !defined(RSI_SIM)
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;\n" ::: "memory");
#endif
}

// order-preserving float <-> unsigned maps for atomicMin / atomicMax on floats
__device__ __forceinline__ u32 f2ord(float f) { u32 u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(u32 u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
}  // namespace rsigpu
