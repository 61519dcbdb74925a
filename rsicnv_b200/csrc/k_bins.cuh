// k_bins.cuh -- per-bin stage on the compacted depth: cap, bin medians, bin sums, chromosome statistics.
//
// Replaces (reference file:line relative to src/):
//   apply_cap (the clamp)              loaddata.cpp:236-238     fused into k_bins' tile load
//   median_transfer                    rsi.cpp:1363-1379        k_bins (warp-per-bin ballot select)
//   negative_binomial_transfer         rsi.cpp:1120-1188        k_bins (bin sums, 31 strided value histograms),
//                                                               k_chr_stats (median, MAD), k_nb_gather / k_nb_scale
//   RDmedian / RDsd in main()          rsi.cpp:2202-2203        k_chr_stats (from the value histograms: exact integer sums)
//
// One pass over the compacted depth (4 B/base read; capped values are written back only where they
// changed).  The 31 strided sub-sample histograms live in one private [class][value] table per warp:
// a warp counts 31 CONSECUTIVE bases per step, which fall into 31 different classes, so the plain
// read-modify-writes of its lanes never collide.
#pragma once
#include "k_load.cuh"

namespace rsigpu {

enum { C_NT = 256, C_NW = C_NT / 32, C_K = 128, C_KP = C_K + 2 /* padded row: 65 words, lanes of a step hit 31 different banks */,
       C_TP = 8192 /* largest bin size */, C_CAP = 13312 /* words of one staged tile (52 KB) */, C_BINS = 128 /* bins per tile at most */ };
// dynamic shared memory: 2 staged tiles | ct[C_NW][MAD_CLASSES][C_KP] u16
#define RSI_SMEM_C (2 * ((size_t)C_CAP * 4 + 16) + (size_t)C_NW * MAD_CLASSES * C_KP * 2)

__device__ __forceinline__ i64 warp_sum_i64(i64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// tile `t` of pass C: bins [b0, b0 + nbt) and their bases [B, B + np); the pseudo-tile after the last bin tile owns the
// < m bases beyond nb*m (they count for the chromosome statistics only)
struct CTile { int b0, nbt, B, np; };
__device__ __forceinline__ CTile c_tile(int t, int nbin_tiles, int bpt, int nb, int m, int Lc) {
  CTile T;
  if (t < nbin_tiles) { T.b0 = t * bpt; T.nbt = imin(bpt, nb - T.b0); T.B = T.b0 * m; T.np = T.nbt * m; }
  else { T.b0 = nb; T.nbt = 0; T.B = nb * m; T.np = Lc - nb * m; }
  return T;
}
// one thread: bulk copy of the tile's words, from the 16-byte aligned address at or below its first base
__device__ __forceinline__ void c_issue(int* dst, u64* bar, const int* __restrict__ rdc, const CTile& T) {
  const int skew = T.B & 3;
  const u32 bytes = (u32)(((T.np + skew) * 4 + 15) & ~15);
  fence_proxy_async();
  mbar_expect_tx(bar, bytes);
  bulk_g2s(dst, rdc + (T.B - skew), bytes, bar);
}

// chist[c * R + v]: #bases of class c (compacted index mod 31, index < 31*floor(Lc/31)) with capped
// value v; thist[v]: the < 31 bases beyond.  (bins_per_tile + 1) * m <= C_CAP - 4.
// One pass over the compacted depth (4 B/base).  Tiles of whole bins arrive by bulk copy into a two-stage ring.  Per tile:
//   phase 1 (all warps): cap clamp + the 31 strided class histograms -- a warp counts 31 CONSECUTIVE bases per step, which
//           fall into 31 different classes, so the plain read-modify-writes of its lanes never collide; the class of a
//           lane is fixed for the whole tile (steps advance by multiples of 31);
//   phase 2: bin medians and sums.  m <= 127: a PAIR of lanes owns a bin, each lane keeps its half of the values in
//           registers as bytes relative to the value window (4 per register) and counts "value <= probe" for 4 values with
//           one subtract (no lane ever waits for another: one shuffle per probe joins the two halves).  Larger m, or a bin
//           with a value outside the window: one warp per bin with ballots.
// tile0 .. tile1: the tiles this launch covers (all of them, or one part of a contig split over several GPUs; the tail
// pseudo-tile has index nbin_tiles and belongs to the last part)
__global__ void __launch_bounds__(C_NT) k_bins(int* __restrict__ rdc, float* __restrict__ bin_med, int* __restrict__ bin_medint,
                                                i64* __restrict__ bin_sum, u32* chist, u32* thist, DevState* st, int bins_per_tile, int tile0, int tile1) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  __shared__ __align__(8) u64 s_bar[2];
  __shared__ int s_slow[C_BINS + 1];
  int* stage[2] = {reinterpret_cast<int*>(smem), reinterpret_cast<int*>(smem + (size_t)C_CAP * 4 + 16)};
  u16* ct = reinterpret_cast<u16*>(smem + 2 * ((size_t)C_CAP * 4 + 16));   // [C_NW][MAD_CLASSES][C_KP]
  const int tid = c.tid, lane = tid & 31, warp = tid >> 5;
  const int Lc = st->Lc, m = st->m, nb = st->nb, R = st->chist_R, cap_on = st->cap_on, capv = st->capv;
  const double thr = st->cap_thr;
  const int sub31 = MAD_CLASSES * (Lc / MAD_CLASSES);
  int wb = 0;
  if (R > C_K) { wb = (int)st->cap_median - C_K / 2; if (wb < 0) wb = 0; if (wb > R - C_K) wb = R - C_K; }
  for (int k = tid; k < C_NW * MAD_CLASSES * C_KP; k += C_NT) ct[k] = 0;
  u16* wt = ct + (size_t)warp * MAD_CLASSES * C_KP;
  const int ithr = thr >= 2147483647.0 ? 0x7fffffff : (int)floor(thr);   // integer v: (double)v > thr  <=>  v > floor(thr)
  i64 mx = 0;
  const int bpt = bins_per_tile;
  const int nbin_tiles = nb > 0 ? (nb + bpt - 1) / bpt : 0;
  const int ntiles = imin(nbin_tiles + (Lc - nb * m > 0 ? 1 : 0), tile1);
  if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
  c.sync();
  if (tid == 0)
    for (int s = 0; s < 2; ++s) { const int t = tile0 + (int)blockIdx.x + s * (int)gridDim.x; if (t < ntiles) c_issue(stage[s], &s_bar[s], rdc, c_tile(t, nbin_tiles, bpt, nb, m, Lc)); }
  const int need = (m - 1) / 2 + 1;                    // median = smallest v with #{x <= v} >= need (m is odd)
  int it = 0;
  for (int tile = tile0 + (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x, ++it) {
    const int s = it & 1;
    const CTile T = c_tile(tile, nbin_tiles, bpt, nb, m, Lc);
    mbar_wait(&s_bar[s], (u32)(it >> 1) & 1u);
#ifdef RSI_SIM_DEBUG
    if (tid == 0 || tid == 255) fprintf(stderr, "k_bins blk %d tid %d tile %d it %d nbt %d np %d m %d\n", (int)blockIdx.x, tid, tile, it, T.nbt, T.np, m);
#endif
    int* vals = stage[s] + (T.B & 3);
    if (tid <= C_BINS) s_slow[tid] = 0;
    // ---- phase 1: clamp + class histograms
    {
      const int cls0 = T.B % MAD_CLASSES;
      int cls = cls0 + lane; if (cls >= MAD_CLASSES) cls -= MAD_CLASSES;       // class of this lane's bases in every step of this tile
      u16* row = wt + cls * C_KP;
      u32* grow = chist + (size_t)cls * R;
      // no two lanes of a warp ever share a row, whatever their relative progress: no warp barrier in the loop
      for (int q = warp * 31 + lane; q < T.np; q += C_NW * 31) {
        if (lane < 31) {
          int v = vals[q];
          if (cap_on && v > ithr) { v = capv; vals[q] = capv; rdc[T.B + q] = capv; }
          const unsigned w = (unsigned)(v - wb);
          if (T.B + q >= sub31) { if (v >= 0 && v < R) atomicAdd(&thist[v], 1u); }
          else if (w < (unsigned)C_K) row[w] += 1;
          else if (v >= 0 && v < R) atomicAdd(&grow[v], 1u);
        }
      }
    }
    c.sync();
    // ---- phase 2: medians and sums
    if (m <= 127) {
      const int nreg = (m + 7) / 8;                   // packed registers per lane: ceil(ceil(m/2) / 4)
      for (int bb = 0; bb < T.nbt; bb += C_NT / 2) {
        const int b = bb + (tid >> 1), par = tid & 1;
        const bool act = b < T.nbt;
        const int* x = vals + (act ? b : 0) * m;
        u32 pk[16];
        int s32 = 0; u32 out = 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          u32 w4 = 0;
          if (r < nreg) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int j = 2 * (r * 4 + k) + par;
              u32 u = 0x7fu;                            // padding: never <= a probe below 127, and a real 127 makes the bin take the slow path
              if (act && j < m) { const int v = x[j]; s32 += v; const unsigned d = (unsigned)(v - wb); out |= d; u = d & 0x7fu; }
              w4 |= u << (8 * k);
            }
          }
          pk[r] = w4;
        }
        // values of the bin must lie in [wb, wb + 126]
        const bool slow = (out >= 127u);
        const i64 sum = (i64)s32 + (i64)__shfl_xor_sync(0xffffffffu, s32, 1);      // < 127 * 2^24
        const int oslow = __shfl_xor_sync(0xffffffffu, (int)slow, 1);               // (no short-circuit around a warp collective)
        const bool pslow = slow || oslow != 0;
        int lo = 0, hi = 126;
#pragma unroll 1
        for (int itb = 0; itb < 7; ++itb) {          // every lane runs the 7 probes (idle pairs on padding): the shuffles stay converged
          const int mid = (lo + hi) >> 1;
          const u32 probe = 0x80808080u | ((u32)mid * 0x01010101u);
          u32 acc = 0;
#pragma unroll
          for (int r = 0; r < 16; ++r) if (r < nreg) acc += ((probe - pk[r]) >> 7) & 0x01010101u;    // per byte: 1 iff value <= mid
          int cnt = (int)((acc * 0x01010101u) >> 24);
          cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
          if (cnt >= need) hi = mid; else lo = mid + 1;
        }
        if (act && par == 0) {
          if (pslow) s_slow[b] = 1;
          else { const int med = wb + lo; bin_med[T.b0 + b] = (float)med; bin_medint[T.b0 + b] = med; bin_sum[T.b0 + b] = sum; mx = lmax(mx, sum); }
        }
      }
      c.sync();
    } else if (tid == 0) { for (int b = 0; b < T.nbt; ++b) s_slow[b] = 1; }
    if (m > 127) c.sync();
    // bins left for the warp-per-bin path (large m, or a value outside the window)
    for (int b = warp; b < T.nbt; b += C_NW) {
      if (!s_slow[b]) continue;
      const int* x = vals + b * m;
      int lo = 0x7fffffff, hi = -0x7fffffff - 1; i64 sm = 0;
      for (int j = lane; j < m; j += 32) { const int v = x[j]; lo = imin(lo, v); hi = imax(hi, v); sm += v; }
      lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
      sm = warp_sum_i64(sm);
      while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        int cnt = 0;
        for (int j0 = 0; j0 < m; j0 += 32) { const int j = j0 + lane; cnt += __popc(__ballot_sync(0xffffffffu, j < m && x[j] <= mid)); }
        if (cnt >= need) hi = mid; else lo = mid + 1;
      }
      if (lane == 0) { bin_med[T.b0 + b] = (float)lo; bin_medint[T.b0 + b] = lo; bin_sum[T.b0 + b] = sm; mx = lmax(mx, sm); }
    }
    c.sync();                                          // everybody is done with stage s
    const int t2 = tile + 2 * (int)gridDim.x;
    if (tid == 0 && t2 < ntiles) c_issue(stage[s], &s_bar[s], rdc, c_tile(t2, nbin_tiles, bpt, nb, m, Lc));
  }
  c.sync();
  for (int item = tid; item < MAD_CLASSES * C_K; item += C_NT) {
    const int cl = item / C_K, w = item % C_K;
    u32 sum = 0;
    for (int k = 0; k < C_NW; ++k) sum += ct[((size_t)k * MAD_CLASSES + cl) * C_KP + w];
    if (sum && wb + w < R) atomicAdd(&chist[cl * R + wb + w], sum);
  }
  mx = c.reduce(mx, MaxOp());
  if (tid == 0 && mx > 0) atomicMax(reinterpret_cast<u64*>(&st->max_binsum), (u64)mx);
}

// ---------------------------------------------------------------------------------------------
// Pass C for bins of m <= 127 bases (the default 101 and every smaller bin): the same work as k_bins with every WARP as an
// independent pipeline, like the per-base passes -- its own warp-tiles of CW_BINS whole bins, its own two-stage bulk-copy
// ring, its own class-histogram table; no block barrier in the loop.  The pseudo-tile after the last bin tile owns the
// < m bases beyond nb*m.
enum { CW_BINS = 16 /* one bin per lane pair */ };
// words of one stage: CW_BINS bins + the skew to the 16-byte aligned source, rounded to 16 bytes (M = 0: any m <= 127)
__host__ __device__ constexpr int cw_cap(int M) { return M ? ((CW_BINS * M + 3 + 3) & ~3) : 2048; }
__host__ __device__ constexpr size_t cw_warp_bytes(int M) { return (size_t)2 * cw_cap(M) * 4 + (((size_t)MAD_CLASSES * C_KP * 2 + 15) & ~(size_t)15); }
// warps per block: as many as the shared memory of an SM holds (a warp's smaller stages leave room for more warps, and more
// warps are what hides the latency of the dependent shared-memory updates)
__host__ __device__ constexpr int cw_nw(int M) { return (int)((size_t)220 * 1024 / cw_warp_bytes(M)) > 16 ? 16 : (int)((size_t)220 * 1024 / cw_warp_bytes(M)); }
// M: the bin size as a compile-time constant (101 = the default, 51), or 0 = read it from the state (any odd m <= 127)
template <int M>
__global__ void __launch_bounds__(cw_nw(M) * 32) k_bins_warp(int* __restrict__ rdc, float* __restrict__ bin_med, int* __restrict__ bin_medint,
                                                      i64* __restrict__ bin_sum, u32* chist, u32* thist, DevState* st, int tile0, int tile1) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  constexpr int CW_NW = cw_nw(M), CW_CAP = cw_cap(M);
  constexpr size_t RSI_CW_WARP_BYTES = cw_warp_bytes(M);
  __shared__ __align__(8) u64 s_bar[CW_NW * 2];
  const int tid = c.tid, lane = tid & 31, warp = tid >> 5;
  unsigned char* wsm = smem + (size_t)warp * RSI_CW_WARP_BYTES;
  u16* wt_ = reinterpret_cast<u16*>(wsm + 2 * CW_CAP * 4);     // [MAD_CLASSES][C_KP]
  u64* bar = s_bar + warp * 2;
  const int Lc = st->Lc, m = M ? M : st->m, nb = st->nb, R = st->chist_R, cap_on = st->cap_on, capv = st->capv;
  const double thr = st->cap_thr;
  const int sub31 = MAD_CLASSES * (Lc / MAD_CLASSES);
  int wb = 0;
  if (R > C_K) { wb = (int)st->cap_median - C_K / 2; if (wb < 0) wb = 0; if (wb > R - C_K) wb = R - C_K; }
  for (int k = lane; k < MAD_CLASSES * C_KP; k += 32) wt_[k] = 0;
  const int ithr = thr >= 2147483647.0 ? 0x7fffffff : (int)floor(thr);   // integer v: (double)v > thr  <=>  v > floor(thr)
  i64 mx = 0;
  const int nbin_tiles = nb > 0 ? (nb + CW_BINS - 1) / CW_BINS : 0;
  const int ntiles = imin(nbin_tiles + (Lc - nb * m > 0 ? 1 : 0), tile1);
  if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  c.sync();
  const int gw = tile0 + (int)blockIdx.x * CW_NW + warp, GW = (int)gridDim.x * CW_NW;
  if (lane == 0)
    for (int s = 0; s < 2; ++s) { const int t = gw + s * GW; if (t < ntiles) c_issue(reinterpret_cast<int*>(wsm + (size_t)s * CW_CAP * 4), &bar[s], rdc, c_tile(t, nbin_tiles, CW_BINS, nb, m, Lc)); }
  const int need = (m - 1) / 2 + 1;                    // median = smallest v with #{x <= v} >= need (m is odd)
  const int nreg = (m + 7) / 8;                        // packed registers per lane: ceil(ceil(m/2) / 4)
  int it = 0;
  for (int tile = gw; tile < ntiles; tile += GW, ++it) {
    const int s = it & 1;
    const CTile T = c_tile(tile, nbin_tiles, CW_BINS, nb, m, Lc);
    int* stage = reinterpret_cast<int*>(wsm + (size_t)s * CW_CAP * 4);
    mbar_wait(&bar[s], (u32)(it >> 1) & 1u);
    int* vals = stage + (T.B & 3);
    // ---- phase 1: cap clamp + class histograms: 31 consecutive bases per step, 31 different classes, a lane's class is fixed for the tile
    if (lane < 31) {
      int cls = T.B % MAD_CLASSES + lane; if (cls >= MAD_CLASSES) cls -= MAD_CLASSES;
      u16* row = wt_ + cls * C_KP;
      u32* grow = chist + (size_t)cls * R;
      const int icap = cap_on ? ithr : 0x7fffffff;
      if (T.B + T.np <= sub31) {            // every tile but the contig's last ones
#pragma unroll 4
        for (int q = lane; q < T.np; q += 31) {
          int v = vals[q];
          if (v > icap) { v = capv; vals[q] = capv; rdc[T.B + q] = capv; }
          const unsigned w = (unsigned)(v - wb);
          if (w < (unsigned)C_K) row[w] += 1;
          else if (v >= 0 && v < R) atomicAdd(&grow[v], 1u);
        }
      } else {
        for (int q = lane; q < T.np; q += 31) {
          int v = vals[q];
          if (v > icap) { v = capv; vals[q] = capv; rdc[T.B + q] = capv; }
          const unsigned w = (unsigned)(v - wb);
          if (T.B + q >= sub31) { if (v >= 0 && v < R) atomicAdd(&thist[v], 1u); }
          else if (w < (unsigned)C_K) row[w] += 1;
          else if (v >= 0 && v < R) atomicAdd(&grow[v], 1u);
        }
      }
    }
    __syncwarp();
    // ---- phase 2: a lane pair per bin; values as bytes relative to the window, 4 per register
    {
      const int b = lane >> 1, par = lane & 1;
      const bool act = b < T.nbt;
      const int* x = vals + (act ? b : 0) * m;
      u32 pk[16];
      int s32 = 0; u32 out = 0;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        u32 w4 = 0;
        if (r < nreg) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = 2 * (r * 4 + k) + par;
            u32 u = 0x7fu;                            // padding: never <= a probe below 127, and a real 127 makes the bin take the slow path
            if (act && j < m) { const int v = x[j]; s32 += v; const unsigned d = (unsigned)(v - wb); out |= d; u = d & 0x7fu; }
            w4 |= u << (8 * k);
          }
        }
        pk[r] = w4;
      }
      const bool slow = (out >= 127u);                // values of the bin must lie in [wb, wb + 126]
      const i64 sum = (i64)s32 + (i64)__shfl_xor_sync(0xffffffffu, s32, 1);      // < 127 * 2^24
      const int oslow = __shfl_xor_sync(0xffffffffu, (int)slow, 1);
      const bool pslow = slow || oslow != 0;
      int lo = 0, hi = 126;
#pragma unroll 1
      for (int itb = 0; itb < 7; ++itb) {            // every lane runs the 7 probes (idle pairs on padding): the shuffles stay converged
        const int mid = (lo + hi) >> 1;
        const u32 probe = 0x80808080u | ((u32)mid * 0x01010101u);
        u32 acc = 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) if (r < nreg) acc += ((probe - pk[r]) >> 7) & 0x01010101u;    // per byte: 1 iff value <= mid
        int cnt = (int)((acc * 0x01010101u) >> 24);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
        if (cnt >= need) hi = mid; else lo = mid + 1;
      }
      if (act && par == 0 && !pslow) { const int med = wb + lo; bin_med[T.b0 + b] = (float)med; bin_medint[T.b0 + b] = med; bin_sum[T.b0 + b] = sum; mx = lmax(mx, sum); }
      // bins with a value outside the window: the whole warp, one bin after the other (rare)
      unsigned todo = __ballot_sync(0xffffffffu, act && par == 0 && pslow);
      while (todo) {
        const int bb = (__ffs((int)todo) - 1) >> 1; todo &= todo - 1;
        const int* xb = vals + bb * m;
        int l2 = 0x7fffffff, h2 = -0x7fffffff - 1; i64 sm = 0;
        for (int j = lane; j < m; j += 32) { const int v = xb[j]; l2 = imin(l2, v); h2 = imax(h2, v); sm += v; }
        l2 = __reduce_min_sync(0xffffffffu, l2); h2 = __reduce_max_sync(0xffffffffu, h2);
        sm = warp_sum_i64(sm);
        while (l2 < h2) {
          const int mid = l2 + ((h2 - l2) >> 1);
          int cnt = 0;
          for (int j0 = 0; j0 < m; j0 += 32) { const int j = j0 + lane; cnt += __popc(__ballot_sync(0xffffffffu, j < m && xb[j] <= mid)); }
          if (cnt >= need) h2 = mid; else l2 = mid + 1;
        }
        if (lane == 0) { bin_med[T.b0 + bb] = (float)l2; bin_medint[T.b0 + bb] = l2; bin_sum[T.b0 + bb] = sm; mx = lmax(mx, sm); }
      }
    }
    __syncwarp();                                      // every lane is done with stage s
    const int t2 = tile + 2 * GW;
    if (lane == 0 && t2 < ntiles) c_issue(stage, &bar[s], rdc, c_tile(t2, nbin_tiles, CW_BINS, nb, m, Lc));
  }
  __syncwarp();
  for (int item = lane; item < MAD_CLASSES * C_K; item += 32) {
    const int cl = item / C_K, w = item % C_K;
    const u32 sum = wt_[cl * C_KP + w];
    if (sum && wb + w < R) atomicAdd(&chist[cl * R + wb + w], sum);
  }
  mx = c.reduce(mx, MaxOp());
  if (tid == 0 && mx > 0) atomicMax(reinterpret_cast<u64*>(&st->max_binsum), (u64)mx);
}

// RDmedian, RDsd (rsi.cpp:2202-2203) and negative_binomial_transfer's MAD (rsi.cpp:1128-1140) from the
// class histograms.  One block; thread c < 31 walks class c.
__global__ void __launch_bounds__(1024) k_chr_stats(const u32* chist, const u32* thist, u32* tot_hist, DevState* st) {
  RSI_CTA_SETUP(c);
  __shared__ double s_mad[MAD_CLASSES];
  const int R = st->chist_R, Lc = st->Lc;
  i64 s1 = 0, s2 = 0;
  for (int v = c.tid; v < R; v += c.nthr) {
    u32 h = thist[v];
    for (int k = 0; k < MAD_CLASSES; ++k) h += chist[k * R + v];
    tot_hist[v] = h;
    s1 += (i64)h * v; s2 += (i64)h * v * v;
  }
  s1 = c.reduce(s1, SumOp()); s2 = c.reduce(s2, SumOp());
  c.sync();
  const u64 n = (u64)Lc;
  int pick[3], fnz, lnz;
  cta_hist_pick(c, tot_hist, R, n / 4, n / 2, n * 3 / 4, pick, &fnz, &lnz);
  const double med = (fnz == lnz || pick[1] < 0) ? (double)s1 / (double)n : (double)pick[1];
  if (c.tid < MAD_CLASSES) {
    // hist-median of int(|x - med|) over the class: first distance d whose cumulative count reaches sub/2
    const u32* h = chist + c.tid * R;
    const int sub = Lc / MAD_CLASSES, imed = (int)med;
    const u64 r2 = (u64)(sub / 2);
    u64 run = 0; int dmin = -1, dmax = -1, hit = -1; double dsum = 0;
    for (int d = 0; d < R; ++d) {
      u64 cnt = 0;
      if (imed + d < R) cnt += h[imed + d];
      if (d > 0 && imed - d >= 0) cnt += h[imed - d];
      if (cnt) { if (dmin < 0) dmin = d; dmax = d; dsum += (double)cnt * d; }
      if (hit < 0 && run < r2 && run + cnt >= r2) hit = d;
      run += cnt;
    }
    double mad = (dmin == dmax || hit < 0) ? (sub > 0 ? dsum / (double)sub : 0.0) : (double)hit;
    s_mad[c.tid] = mad;
  }
  c.sync();
  if (c.tid == 0) {
    st->rdmedian = med;
    const double mean = (double)s1 / (double)n;
    st->rdsd = sqrt((double)s2 / (double)n - mean * mean);
    // hist-median (dy = 0.01) of the 31 class MADs (rsi.cpp:1138)
    double ymin = s_mad[0], ymax = s_mad[0], sum = 0;
    for (int k = 0; k < MAD_CLASSES; ++k) { sum += s_mad[k]; ymin = s_mad[k] < ymin ? s_mad[k] : ymin; ymax = s_mad[k] > ymax ? s_mad[k] : ymax; }
    double mad = sum / (double)MAD_CLASSES;
    if (!((ymax - ymin) < 0.01)) {
      const u64 np = (u64)((ymax - ymin) / 0.01 + 2);
      const u64 i2 = MAD_CLASSES / 2;
      // walk the occupied buckets in increasing order (31 values: selection by repeated minimum)
      u64 run = 0, prev = 0; bool have_prev = false;
      for (;;) {
        u64 best = ~0ull;
        for (int k = 0; k < MAD_CLASSES; ++k) {
          const u64 b = (u64)((s_mad[k] - ymin) / 0.01 + 0.5);
          if ((!have_prev || b > prev) && b < best) best = b;
        }
        if (best == ~0ull || best >= np) break;
        u64 cnt = 0;
        for (int k = 0; k < MAD_CLASSES; ++k) if ((u64)((s_mad[k] - ymin) / 0.01 + 0.5) == best) ++cnt;
        if (run < i2 && run + cnt >= i2) { mad = ymin + (double)best * 0.01; break; }
        run += cnt; prev = best; have_prev = true;
      }
    }
    st->rdmad = mad;
    if (!(mad > 0) || !(med > 0)) st->err |= ERR_DEGENERATE;
    st->lim_del = med * 0.75; st->lim_dup = med * 1.25;
  }
}

// Negative-binomial transform, step 1: gather the host-built table (glibc log/sqrt: bit-exact by
// construction, SURVEY hard part 3) by bin sum, track the minimum.
__global__ void k_nb_gather(const i64* __restrict__ bin_sum, const float* __restrict__ lut, int lut_n, float* __restrict__ out, DevState* st) {
  RSI_CTA_SETUP(c);
  const int nb = st->nb;
  u32 mn = 0xffffffffu;
  for (int b = (int)(blockIdx.x * blockDim.x + threadIdx.x); b < nb; b += (int)(gridDim.x * blockDim.x)) {
    i64 s = bin_sum[b];
    if (s >= lut_n) s = lut_n - 1;
    const float t = lut[s];
    out[b] = t;
    mn = f2ord(t) < mn ? f2ord(t) : mn;
  }
  mn = c.reduce(mn, MinOp());
  if (c.tid == 0) atomicMin(&st->nb_tmin_ord, mn);
}
// step 2: subtract the minimum, rescale so that t(median * m) -> median, write the anchors into
// bins 0..2 (rsi.cpp:1165-1185).
__global__ void k_nb_scale(float* __restrict__ out, DevState* st) {
  const int nb = st->nb;
  const double tmin = (double)ord2f(st->nb_tmin_ord);
  const double medn = st->med_nbt_raw - tmin;
  const double rdm = st->rdmedian;
  for (int b = (int)(blockIdx.x * blockDim.x + threadIdx.x); b < nb; b += (int)(gridDim.x * blockDim.x)) {
    float t = (float)((double)out[b] - tmin);
    t = (float)__dmul_rn(__ddiv_rn((double)t, medn), rdm);
    if (b == 0) t = (float)__dmul_rn(__ddiv_rn(st->del_nbt_raw - tmin, medn), rdm);
    if (b == 1) t = (float)__dmul_rn(__ddiv_rn(st->dup_nbt_raw - tmin, medn), rdm);
    if (b == 2) t = (float)__dmul_rn(__ddiv_rn(medn, medn), rdm);
    out[b] = t;
  }
}

}  // namespace rsigpu
