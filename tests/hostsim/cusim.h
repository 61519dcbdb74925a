// cusim.h -- TEST-ONLY single-process emulation of the CUDA execution model (x86-64 Linux).
//
// There is no GPU in the development container, so the CUDA sources of rsicnv_b200/csrc are ALSO
// compiled with g++ against this header (-DRSI_SIM) into tests/hostsim/librsigpu_sim.so, which the
// CPU test-suite drives through the same C-ABI as the real library.  Every thread of a block is a
// fiber; blocks run one after the other; __syncthreads and the warp collectives are scheduling
// points.  CUSIM_ORDER=0/1/2 (forward / reverse / rotating) changes the order in which the fibers
// of a block are resumed so that a missing barrier shows up as a parity failure in at least one
// order.  Nothing here is ever linked into the product library (rsicnv_b200/librsigpu.so is built by
// nvcc only and has no CPU path).
#pragma once
#if !defined(__x86_64__)
#error "cusim.h supports x86-64 only"
#endif
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct float4 { float x, y, z, w; };
inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline int2 make_int2(int a, int b) { return int2{a, b}; }

namespace cusim {

struct Warp {
  alignas(16) unsigned char slot[32][16];
  alignas(16) unsigned char result[32][16];
  unsigned arrived = 0, ready = 0, alive = 0;
};
struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  bool done = false;
  unsigned long long bar_gen_wait = 0;
};
struct BlockCtx {     // one resident thread block (a cluster launch keeps several resident at once)
  std::vector<Warp> warps;
  int live = 0;
  unsigned long long bar_gen = 0; int bar_arrived = 0; int bar_acc_or = 0, bar_acc_and = 1, bar_acc_cnt = 0;
  int bar_res_or[2] = {0, 0}, bar_res_and[2] = {1, 1}, bar_res_cnt[2] = {0, 0};
  unsigned char* dyn_smem = nullptr;
  uint3 bidx{0, 0, 0};
};
struct State {
  std::vector<Fiber> fib;        // [block-in-cluster][thread] flattened
  std::vector<BlockCtx> blk;
  int nthreads = 0, nblk = 1, cur = -1, live = 0;
  unsigned long long cl_gen = 0; int cl_arrived = 0;   // cluster barrier
  void* sched_sp = nullptr;
  std::function<void()> body;
  unsigned long long progress = 0;
  dim3 block_dim, grid_dim;
  BlockCtx& cb() { return blk[cur / nthreads]; }
  int ct() const { return cur % nthreads; }
  unsigned char* dyn() { return cb().dyn_smem; }
};
inline State& S() { static State s; return s; }

inline uint3& tIdx() { static uint3 v; return v; }
inline uint3& bIdx() { static uint3 v; return v; }
inline dim3& bDim() { static dim3 v; return v; }
inline dim3& gDim() { static dim3 v; return v; }

// ---- context switch: save callee-saved registers on the current stack, swap stack pointers
extern "C" void cusim_switch(void** save_sp, void* new_sp);
#if defined(CUSIM_IMPL)
asm(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size cusim_switch,.-cusim_switch
)");
#endif

inline void yield_to_sched() {
  State& s = S();
  Fiber& f = s.fib[s.cur];
  cusim_switch(&f.sp, s.sched_sp);
}
inline void release_block_barrier(BlockCtx& b) {
  int slot = (int)(b.bar_gen & 1);
  b.bar_res_or[slot] = b.bar_acc_or; b.bar_res_and[slot] = b.bar_acc_and; b.bar_res_cnt[slot] = b.bar_acc_cnt;
  b.bar_acc_or = 0; b.bar_acc_and = 1; b.bar_acc_cnt = 0; b.bar_arrived = 0; b.bar_gen++;
}
void fiber_entry();
#if defined(CUSIM_IMPL)
void fiber_entry() {
  State& s = S();
  s.body();
  Fiber& f = s.fib[s.cur];
  BlockCtx& b = s.cb();
  f.done = true;
  s.live--; b.live--;
  b.warps[s.ct() >> 5].alive &= ~(1u << (s.ct() & 31));
  s.progress++;
  // a thread that exits releases a barrier the others are waiting on
  if (b.live > 0 && b.bar_arrived == b.live) release_block_barrier(b);
  if (s.live > 0 && s.cl_arrived == s.live) { s.cl_arrived = 0; s.cl_gen++; }
  cusim_switch(&f.sp, s.sched_sp);
  std::abort();
}
#endif

static const size_t kStack = 96 * 1024;

inline void set_indices(int f) {
  State& s = S();
  const int t = f % s.nthreads;
  uint3& ti = tIdx();
  ti.x = t % s.block_dim.x; ti.y = (t / s.block_dim.x) % s.block_dim.y; ti.z = t / (s.block_dim.x * s.block_dim.y);
  bIdx() = s.blk[f / s.nthreads].bidx;
}

void run_resident(int order_mode);
#if defined(CUSIM_IMPL)
void run_resident(int order_mode) {   // runs the s.nblk resident blocks (1, or a whole cluster) to completion
  State& s = S();
  const int n = s.nthreads, nf = n * s.nblk;
  s.live = nf; s.cl_gen = 0; s.cl_arrived = 0;
  const int nw = (n + 31) / 32;
  for (int b = 0; b < s.nblk; ++b) {
    BlockCtx& B = s.blk[b];
    B.live = n; B.bar_gen = 0; B.bar_arrived = 0; B.bar_acc_or = 0; B.bar_acc_and = 1; B.bar_acc_cnt = 0;
    B.warps.assign(nw, Warp());
    for (int t = 0; t < n; ++t) B.warps[t >> 5].alive |= 1u << (t & 31);
  }
  for (int t = 0; t < nf; ++t) {
    Fiber& f = s.fib[t];
    f.done = false; f.bar_gen_wait = 0;
    // initial frame: six zeroed callee-saved registers, then the entry address for `ret`
    uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // fake return address slot (keeps rsp%16==8 at entry)
    *--sp = (void*)&fiber_entry;
    for (int k = 0; k < 6; ++k) *--sp = nullptr;
    f.sp = (void*)sp;
  }
  unsigned long long rounds = 0;
  while (s.live > 0) {
    unsigned long long before = s.progress;
    for (int k = 0; k < nf; ++k) {
      int t = k;
      if (order_mode == 1) t = nf - 1 - k;
      else if (order_mode == 2) t = (int)((k + rounds * 7) % (unsigned long long)nf);
      Fiber& f = s.fib[t];
      if (f.done) continue;
      s.cur = t;
      set_indices(t);
      cusim_switch(&s.sched_sp, f.sp);
    }
    ++rounds;
    if (s.progress == before) {
      std::fprintf(stderr, "cusim: deadlock (divergent barrier / collective) near block (%u,%u,%u)\n", bIdx().x, bIdx().y, bIdx().z);
      std::abort();
    }
  }
}
#endif

inline int order_mode() {
  static int m = -1;
  if (m < 0) { const char* e = std::getenv("CUSIM_ORDER"); m = e ? std::atoi(e) : 0; }
  return m;
}

void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body, int cluster = 1);
#if defined(CUSIM_IMPL)
void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body, int cluster) {
  State& s = S();
  if (s.cur >= 0 && s.live > 0) { std::fprintf(stderr, "cusim: nested launch\n"); std::abort(); }
  const int n = (int)(block.x * block.y * block.z);
  static int trace = -1, seq = 0;
  if (trace < 0) trace = std::getenv("CUSIM_TRACE") ? 1 : 0;
  if (trace) std::fprintf(stderr, "cusim: launch #%d grid %u block %d smem %zu cluster %d\n", seq++, grid.x, n, smem, cluster);
  if (n <= 0 || n > 1024) { std::fprintf(stderr, "cusim: bad block size %d\n", n); std::abort(); }
  if (cluster < 1 || grid.x % (unsigned)cluster) { std::fprintf(stderr, "cusim: grid.x not a multiple of the cluster size\n"); std::abort(); }
  const int nf = n * cluster;
  if ((int)s.fib.size() < nf) {
    size_t old = s.fib.size();
    s.fib.resize(nf);
    for (size_t t = old; t < (size_t)nf; ++t) s.fib[t].stack = (char*)std::malloc(kStack);
  }
  std::vector<std::vector<unsigned char>> dyn(cluster, std::vector<unsigned char>(smem + 64));
  s.nthreads = n; s.nblk = cluster; s.block_dim = block; s.grid_dim = grid; s.body = std::move(body);
  s.blk.assign(cluster, BlockCtx());
  bDim() = block; gDim() = grid;
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; x += (unsigned)cluster) {
        for (int b = 0; b < cluster; ++b) {
          std::memset(dyn[b].data(), 0xCD, dyn[b].size());  // shared memory is NOT zero-initialised on a GPU
          s.blk[b].dyn_smem = (unsigned char*)(((uintptr_t)dyn[b].data() + 15) & ~(uintptr_t)15);
          s.blk[b].bidx = uint3{x + (unsigned)b, y, z};
        }
        run_resident(order_mode());
      }
  s.cur = -1; s.live = 0;
}
#endif

// ---- cluster: rank of the block, size, barrier over every thread of every block of the cluster
inline unsigned cluster_ctarank() { return (unsigned)(S().cur / S().nthreads); }
inline unsigned cluster_nctarank() { return (unsigned)S().nblk; }
inline void cluster_sync() {
  State& s = S();
  const unsigned long long gen = s.cl_gen;
  s.cl_arrived++;
  s.progress++;
  if (s.cl_arrived == s.live) { s.cl_arrived = 0; s.cl_gen++; }
  else while (s.cl_gen == gen) yield_to_sched();
}

// ---- block barrier
inline int barrier(int pred, int kind) {  // kind 0 plain, 1 or, 2 and, 3 count
  State& s = S();
  BlockCtx& b = s.cb();
  Fiber& f = s.fib[s.cur];
  b.bar_acc_or |= (pred != 0); b.bar_acc_and &= (pred != 0); b.bar_acc_cnt += (pred != 0);
  const unsigned long long gen = b.bar_gen;
  b.bar_arrived++;
  s.progress++;
  if (b.bar_arrived == b.live) release_block_barrier(b);
  else {
    f.bar_gen_wait = gen;
    while (b.bar_gen == gen) yield_to_sched();
  }
  int slot = (int)(gen & 1);
  return kind == 1 ? b.bar_res_or[slot] : kind == 2 ? b.bar_res_and[slot] : kind == 3 ? b.bar_res_cnt[slot] : 0;
}

// ---- warp collective: the last lane to arrive computes every participant's result
template <class T, class R, class F>
inline R warp_coll(unsigned mask, const T& val, F f) {
  static_assert(sizeof(T) <= 16 && sizeof(R) <= 16, "payload");
  State& s = S();
  Warp& W = s.cb().warps[s.ct() >> 5];
  const int lane = s.ct() & 31;
  const unsigned bit = 1u << lane;
  mask &= W.alive;
  if (!(mask & bit)) { std::fprintf(stderr, "cusim: lane %d not in its own mask\n", lane); std::abort(); }
  std::memcpy(W.slot[lane], &val, sizeof(T));
  W.arrived |= bit;
  s.progress++;
  if ((W.arrived & mask) == mask) {
    T vals[32];
    for (int l = 0; l < 32; ++l) std::memcpy(&vals[l], W.slot[l], sizeof(T));
    for (int l = 0; l < 32; ++l)
      if (mask & (1u << l)) { R r = f(l, vals, mask); std::memcpy(W.result[l], &r, sizeof(R)); }
    W.ready |= mask;
    W.arrived &= ~mask;
  } else {
    while (!(W.ready & bit)) yield_to_sched();
  }
  W.ready &= ~bit;
  R r;
  std::memcpy(&r, W.result[lane], sizeof(R));
  return r;
}

}  // namespace cusim

#define threadIdx (cusim::tIdx())
#define blockIdx (cusim::bIdx())
#define blockDim (cusim::bDim())
#define gridDim (cusim::gDim())
static const int warpSize = 32;

inline void __syncthreads() { cusim::barrier(0, 0); }
inline int __syncthreads_or(int p) { return cusim::barrier(p, 1); }
inline int __syncthreads_and(int p) { return cusim::barrier(p, 2); }
inline int __syncthreads_count(int p) { return cusim::barrier(p, 3); }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  cusim::warp_coll<int, int>(mask, 0, [](int, const int*, unsigned) { return 0; });
}
inline void __threadfence() {}
inline void __threadfence_block() {}
inline unsigned __activemask() { return cusim::S().cb().warps[cusim::S().ct() >> 5].alive; }

// every lane may name a different source lane: the source index travels with the value
template <class T> struct CusimShflArg { T v; int src; };
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  CusimShflArg<T> arg; arg.v = v; arg.src = src;
  return cusim::warp_coll<CusimShflArg<T>, T>(mask, arg, [=](int l, const CusimShflArg<T>* a, unsigned) { int base = l & ~(width - 1); return a[base + (a[l].src & (width - 1))].v; });
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lm, int width = 32) {
  return cusim::warp_coll<T, T>(mask, v, [=](int l, const T* a, unsigned) { int o = l ^ lm; return (o / width == l / width && o < 32) ? a[o] : a[l]; });
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
  return cusim::warp_coll<T, T>(mask, v, [=](int l, const T* a, unsigned) { int o = l - (int)d; return (o >= (l & ~(width - 1))) ? a[o] : a[l]; });
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
  return cusim::warp_coll<T, T>(mask, v, [=](int l, const T* a, unsigned) { int o = l + (int)d; return (o < (l & ~(width - 1)) + width) ? a[o] : a[l]; });
}
inline unsigned __ballot_sync(unsigned mask, int p) {
  return cusim::warp_coll<int, unsigned>(mask, p, [](int, const int* a, unsigned m) { unsigned r = 0; for (int l = 0; l < 32; ++l) if ((m >> l & 1) && a[l]) r |= 1u << l; return r; });
}
inline int __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0; }
inline int __all_sync(unsigned mask, int p) {
  return cusim::warp_coll<int, int>(mask, p, [](int, const int* a, unsigned m) { for (int l = 0; l < 32; ++l) if ((m >> l & 1) && !a[l]) return 0; return 1; });
}
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
  return cusim::warp_coll<T, unsigned>(mask, v, [](int me, const T* a, unsigned m) { unsigned r = 0; for (int l = 0; l < 32; ++l) if ((m >> l & 1) && a[l] == a[me]) r |= 1u << l; return r; });
}
template <class T> inline T __reduce_add_sync(unsigned mask, T v) {
  return cusim::warp_coll<T, T>(mask, v, [](int, const T* a, unsigned m) { T r = 0; for (int l = 0; l < 32; ++l) if (m >> l & 1) r += a[l]; return r; });
}
template <class T> inline T __reduce_min_sync(unsigned mask, T v) {
  return cusim::warp_coll<T, T>(mask, v, [](int me, const T* a, unsigned m) { T r = a[me]; for (int l = 0; l < 32; ++l) if ((m >> l & 1) && a[l] < r) r = a[l]; return r; });
}
template <class T> inline T __reduce_max_sync(unsigned mask, T v) {
  return cusim::warp_coll<T, T>(mask, v, [](int me, const T* a, unsigned m) { T r = a[me]; for (int l = 0; l < 32; ++l) if ((m >> l & 1) && a[l] > r) r = a[l]; return r; });
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  return cusim::warp_coll<unsigned, unsigned>(mask, v, [](int, const unsigned* a, unsigned m) { unsigned r = 0; for (int l = 0; l < 32; ++l) if (m >> l & 1) r |= a[l]; return r; });
}

// ---- atomics (one fiber runs at a time)
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---- intrinsics
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31u; return sh ? (lo >> sh) | (hi << (32u - sh)) : lo; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (v >> i & 1) r |= 1u << (31 - i); return r; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {
  unsigned long long v = ((unsigned long long)y << 32) | x; unsigned r = 0;
  for (int i = 0; i < 4; ++i) { unsigned n = (sel >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * n)) & 0xff) << (8 * i); }
  return r;
}
inline unsigned __vcmpgeu4(unsigned a, unsigned b) {
  unsigned m = 0;
  for (int k = 0; k < 4; ++k) if (((a >> (8 * k)) & 0xff) >= ((b >> (8 * k)) & 0xff)) m |= 0xffu << (8 * k);
  return m;
}
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline double __ll2double_rn(long long v) { return (double)v; }
inline float __double2float_rn(double v) { return (float)v; }
using std::max;
using std::min;

// ---- the slice of the runtime API the host side of librsigpu uses
typedef int cudaError_t;
typedef void* cudaStream_t;
struct CusimEvent { std::chrono::steady_clock::time_point t; };
typedef CusimEvent* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyHostToHost = 0, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; char name[64]; };
inline const char* cudaGetErrorString(cudaError_t) { return "cusim error"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; std::strcpy(p->name, "cusim"); return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); if (!*p) return cudaErrorMemoryAllocation; std::memset(*p, 0xA5, n); return 0; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 0; return 0; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return 0; }
inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t = nullptr) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new CusimEvent(); return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return 0; }
enum { cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
