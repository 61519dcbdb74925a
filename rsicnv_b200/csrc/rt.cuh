// rt.cuh -- the one place that knows whether the sources are compiled by nvcc for sm_100a (the
// product) or by g++ against tests/hostsim/cusim.h (RSI_SIM: the CPU emulator the test-suite uses
// because the development container has no GPU).  The product library never defines RSI_SIM.
#pragma once
#include <stdint.h>
#include <string.h>

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

// ---- TMA-style staging: 1-D bulk copies global -> shared memory (cp.async.bulk, SASS UBLKCP) completing on an mbarrier.
// One elected thread arms the barrier with the byte count and issues the copies; every thread of the block then waits on the
// barrier's phase.  Tiles stay in flight independent of occupancy (no registers are held by outstanding loads), which is
// what the per-base passes need: 8 warps per SM with two 36 KB tiles in flight instead of 8 loads per thread and a barrier.
// Requirements of the instruction: 16-byte aligned source and destination, size a multiple of 16.
#if defined(RSI_SIM)
// emulator: the copy happens at issue; a barrier word = completed phases (low 32 bits) | pending bytes (high 32 bits)
__device__ __forceinline__ void mbar_init(u64* bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) { *bar += (u64)bytes << 32; }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  memcpy(dst, src, bytes);
  *bar -= (u64)bytes << 32;
  if ((*bar >> 32) == 0) { *bar += 1; cusim::S().progress++; }
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) { while (((u32)*bar & 1u) == parity) cusim::yield_to_sched(); }
__device__ __forceinline__ void fence_proxy_async() {}
#else
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// generic-proxy accesses to a shared buffer (ours) ordered before the async proxy (the next bulk copy) writes it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// order-preserving float <-> unsigned maps for atomicMin / atomicMax on floats
__device__ __forceinline__ u32 f2ord(float f) { u32 u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(u32 u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
}  // namespace rsigpu
