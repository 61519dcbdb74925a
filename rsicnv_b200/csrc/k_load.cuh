// k_load.cuh -- per-base streaming stage: raw depth -> GC-adjusted, capped, N-compacted depth.
//
// Replaces (reference file:line relative to src/):
//   checkgccontent / adjustgccontent   gccontent.cpp:95-184, 43-92     k_gc_table, k_gc_finalize, k_gc_adjust
//   apply_cap                          loaddata.cpp:229-240            k_cap_params (+ clamp fused into k_bins)
//   concatenate_data                   loaddata.cpp:48-85              fused into k_gc_adjust's store
//   get_N_regions                      readref.cpp:88-112              k_n_runs
//
// Layout in HBM: raw depth int32[L] and FASTA bytes[L] (both padded to a multiple of 16 bytes and a
// further 256 bytes so vector loads never leave the allocation); the compacted depth int32[Lc].
// The histogram-type accumulations (GC strata table, value histogram) use one PRIVATE table per WARP in
// shared memory: a warp handles 32 consecutive bases per step, lanes with the same key are merged with
// __match_any_sync / __reduce_add_sync and the group leader does a plain read-modify-write -- no atomics
// (neighbouring bases have nearly equal keys, so shared/global atomics would serialise).  Tables are
// flushed once per block at the end of a persistent loop over tiles.
#pragma once
#include "cta.cuh"
#include "state.cuh"

namespace rsigpu {

#define RSI_CTA_SETUP(c)                                   \
  __shared__ __align__(16) unsigned char cta_red_[33 * 16]; \
  __shared__ double cta_bc_[32];                            \
  Cta c;                                                   \
  c.tid = (int)threadIdx.x; c.nthr = (int)blockDim.x; c.red = cta_red_; c.bc = cta_bc_;

enum { LD_TILE = 4096, LD_FAB = LD_TILE + 240, LD_PRE = LD_TILE + 208 };
enum { A_NT = 512, A_NW = A_NT / 32, B_NT = 512, B_NW = B_NT / 32, B_K = 256 };

// ---------------------------------------------------------------------------------------------
// N runs of the contig (uppercase 'N' only): run starts and run ends are appended (unordered) to two
// lists; the host sorts the few hundred entries, pads and merges them (get_noseq_regions,
// loaddata.cpp:243-273).
__global__ void k_n_runs(const u8* __restrict__ fa, int L, int* beg, int* end, int* n_beg, int* n_end, int cap) {
  for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < L; i += (int)(gridDim.x * blockDim.x)) {
    if (fa[i] != 'N') continue;
    if (i == 0 || fa[i - 1] != 'N') { int k = atomicAdd(n_beg, 1); if (k < cap) beg[k] = i; }
    if (i == L - 1 || fa[i + 1] != 'N') { int k = atomicAdd(n_end, 1); if (k < cap) end[k] = i; }
  }
}

// ---------------------------------------------------------------------------------------------
// GC-window prefix of one tile.  The 201-bp window of base i is [lo, lo+200] with
// lo = clamp(i-100, 0, L-202) (gccontent.cpp:124-132 incl. the "right edge never takes the last
// base" quirk).  After the call pre[k] = #GC in fasta[wlo, wlo+k), and
// nGC(i) = pre[lo(i)-wlo+201] - pre[lo(i)-wlo].
__device__ __forceinline__ int gc_lo(int i, int L) { return iclamp(i - GC_WIN / 2, 0, L - GC_WIN - 1); }

__device__ int tile_gc_prefix(const Cta& c, const u8* __restrict__ fa, int L, int t0, int t1, u8* fab, u16* pre) {
  const int wlo = gc_lo(t0, L), whi = gc_lo(t1 - 1, L) + GC_WIN;
  const int nf = whi - wlo;
  const int a0 = wlo & ~15;
  const int nvec = (whi - a0 + 15) >> 4;
  const uint4* src = reinterpret_cast<const uint4*>(fa + a0);
  uint4* dst = reinterpret_cast<uint4*>(fab);
  for (int v = c.tid; v < nvec; v += c.nthr) dst[v] = src[v];
  c.sync();
  const int off = wlo - a0;
  const int chunk = (nf + c.nthr - 1) / c.nthr;
  const int k0 = imin(c.tid * chunk, nf), k1 = imin(k0 + chunk, nf);
  int local = 0;
  for (int k = k0; k < k1; ++k) { u8 ch = fab[off + k]; local += (ch == 'G' || ch == 'C') ? 1 : 0; }
  int tot;
  int run = c.scan_excl(local, &tot);
  for (int k = k0; k < k1; ++k) { pre[k] = (u16)run; u8 ch = fab[off + k]; run += (ch == 'G' || ch == 'C') ? 1 : 0; }
  if (c.tid == 0) pre[nf] = (u16)tot;
  c.sync();
  return wlo;
}

// ---------------------------------------------------------------------------------------------
// Pass A: mean of the positive depths and the per-stratum depth sums / counts (5 B/base read).
// Dynamic shared memory: tsum[A_NW][GC_STRATA] u64 | tcnt[A_NW][GC_STRATA] u32 | fab | pre
#define RSI_SMEM_A ((size_t)A_NW * GC_STRATA * 12 + LD_FAB + (LD_PRE + 8) * 2)
__global__ void __launch_bounds__(A_NT) k_gc_table(const int* __restrict__ rd, const u8* __restrict__ fa, DevState* st) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  const int L = st->L, do_gc = st->gc_on;
  u64* tsum = reinterpret_cast<u64*>(smem);
  u32* tcnt = reinterpret_cast<u32*>(tsum + A_NW * GC_STRATA);
  u8* fab = reinterpret_cast<u8*>(tcnt + A_NW * GC_STRATA);
  u16* pre = reinterpret_cast<u16*>(fab + LD_FAB);
  const int tid = c.tid, lane = tid & 31, warp = tid >> 5;
  if (do_gc) for (int k = tid; k < A_NW * GC_STRATA; k += A_NT) { tsum[k] = 0ull; tcnt[k] = 0u; }
  u64* wsum = tsum + warp * GC_STRATA; u32* wcnt = tcnt + warp * GC_STRATA;
  u64 psum = 0, pcnt = 0;
  int vmin = 0x7fffffff, vmax = -0x7fffffff - 1;
  const int ntiles = (L + LD_TILE - 1) / LD_TILE;
  for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x) {
    const int t0 = tile * LD_TILE, t1 = imin(t0 + LD_TILE, L);
    int v[LD_TILE / A_NT];
#pragma unroll
    for (int j = 0; j < LD_TILE / A_NT; ++j) v[j] = rd[t0 + warp * (LD_TILE / A_NW) + j * 32 + lane];   // a warp owns 256 consecutive bases
    int wlo = 0;
    if (do_gc) { c.sync(); wlo = tile_gc_prefix(c, fa, L, t0, t1, fab, pre); }
#pragma unroll
    for (int j = 0; j < LD_TILE / A_NT; ++j) {
      const int p = t0 + warp * (LD_TILE / A_NW) + j * 32 + lane;
      const bool valid = p < t1;
      const int x = v[j];
      if (valid) { vmin = imin(vmin, x); vmax = imax(vmax, x); if (x > 0) { psum += (u64)x; pcnt += 1; } }
      if (do_gc) {
        int g = 0x10000 + lane;    // invalid lanes: a key nobody shares
        if (valid) { const int lo = gc_lo(p, L) - wlo; g = (int)pre[lo + GC_WIN] - (int)pre[lo]; }
        const unsigned m = __match_any_sync(0xffffffffu, g);
        const int sum = __reduce_add_sync(m, valid ? (x & 0xffffff) : 0);
        if (valid && lane == __ffs((int)m) - 1) { wsum[g] += (u64)(u32)sum; wcnt[g] += (u32)__popc(m); }
        __syncwarp();
      }
    }
  }
  c.sync();
  if (do_gc) {
    for (int g = tid; g < GC_STRATA; g += A_NT) {
      u64 s = 0, n = 0;
      for (int w = 0; w < A_NW; ++w) { s += tsum[w * GC_STRATA + g]; n += tcnt[w * GC_STRATA + g]; }
      if (n) { atomicAdd(&st->gc_sum[g], s); atomicAdd(&st->gc_cnt[g], n); }
    }
  }
  psum = c.reduce(psum, SumOp()); pcnt = c.reduce(pcnt, SumOp());
  vmin = c.reduce(vmin, MinOp()); vmax = c.reduce(vmax, MaxOp());
  if (tid == 0) {
    atomicAdd(&st->pos_sum, psum); atomicAdd(&st->pos_cnt, pcnt);
    atomicMin(&st->rd_min, vmin); atomicMax(&st->rd_max, vmax);
  }
}

// RDmean, the stratum table (empty or <1 -> RDmean), the tail-quirk constants and the window base of
// pass B's private histogram.  One block of GC_STRATA+ threads.
__global__ void k_gc_finalize(const u8* __restrict__ fa, DevState* st) {
  RSI_CTA_SETUP(c);
  const int L = st->L;
  double mean = 0.0;
  if (st->pos_cnt > 0) mean = (double)st->pos_sum / (double)st->pos_cnt;
  for (int g = c.tid; g < GC_STRATA; g += c.nthr) {
    double t = mean;
    if (st->gc_cnt[g] > 0) t = (double)st->gc_sum[g] / (double)st->gc_cnt[g];
    if (t < 1) t = mean;
    st->gc_tab[g] = t;
  }
  int local = 0;
  for (int k = L - GC_WIN + c.tid; k < L; k += c.nthr) { u8 ch = fa[k]; local += (ch == 'G' || ch == 'C') ? 1 : 0; }
  local = c.reduce(local, SumOp());
  if (c.tid == 0) {
    st->rdmean = mean;
    st->gstar = local;
    st->s20 = 20 * (L / 20);
    st->r20 = L - 20 * (L / 20);
    int hb = (int)mean - B_K / 2;
    st->hist_base = hb < 0 ? 0 : hb;
    if (st->rd_min < 0 || st->rd_max >= (1 << 24)) st->err |= ERR_DEPTH_RANGE;
  }
}

// ---------------------------------------------------------------------------------------------
// Pass B: GC adjust (out-of-place map + the 21st pseudo-slice quirk), value histogram of
// ALL positions for apply_cap's median, and the N-compacted store (9 B/base: 4+1 read, 4 written).
// noseq intervals: nbeg/nend (0-based inclusive), ncum[k] = bases removed by intervals 0..k-1.
// Dynamic shared memory: vh[B_NW][B_K] u32 | tab[GC_STRATA] f64 | fab | pre
#define RSI_SMEM_B ((size_t)B_NW * B_K * 4 + GC_STRATA * 8 + LD_FAB + (LD_PRE + 8) * 2)
__global__ void __launch_bounds__(B_NT) k_gc_adjust(const int* __restrict__ rd, const u8* __restrict__ fa, int* __restrict__ rdc,
                                                     const int* __restrict__ nbeg, const int* __restrict__ nend, const int* __restrict__ ncum,
                                                     u32* hist_all, DevState* st) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  __shared__ int s_k0, s_hasn;
  const int L = st->L, do_gc = st->gc_on, nn = st->n_noseq;
  u32* vh = reinterpret_cast<u32*>(smem);                          // [B_NW][B_K]
  double* tab = reinterpret_cast<double*>(smem + (size_t)B_NW * B_K * 4);
  u8* fab = reinterpret_cast<u8*>(tab + GC_STRATA);
  u16* pre = reinterpret_cast<u16*>(fab + LD_FAB);
  const int tid = c.tid, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < B_NW * B_K; k += B_NT) vh[k] = 0;
  for (int g = tid; g < GC_STRATA; g += B_NT) tab[g] = st->gc_tab[g];
  u32* wh = vh + warp * B_K;
  const double mean = st->rdmean;
  const int hb = st->hist_base, s20 = st->s20, r20 = st->r20, gstar = st->gstar;
  const int q0 = s20 + r20 - GC_WIN;   // first overwritten position of the pseudo-slice (r20 >= 2)
  int bad = 0;
  u32 zeros = 0;
  const int ntiles = (L + LD_TILE - 1) / LD_TILE;
  for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x) {
    const int t0 = tile * LD_TILE, t1 = imin(t0 + LD_TILE, L);
    int v[LD_TILE / B_NT];
#pragma unroll
    for (int j = 0; j < LD_TILE / B_NT; ++j) v[j] = rd[t0 + warp * (LD_TILE / B_NW) + j * 32 + lane];
    c.sync();
    if (tid == 0) {  // first interval that ends at or after t0
      int lo = 0, hi = nn;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (nend[mid] < t0) lo = mid + 1; else hi = mid; }
      s_k0 = lo; s_hasn = (lo < nn && nbeg[lo] < t1) ? 1 : 0;
    }
    int wlo = 0;
    if (do_gc) wlo = tile_gc_prefix(c, fa, L, t0, t1, fab, pre); else c.sync();
    const int k0 = s_k0, hasn = s_hasn;
    const int shift0 = k0 < nn ? ncum[k0] : (nn ? ncum[nn - 1] + (nend[nn - 1] - nbeg[nn - 1] + 1) : 0);
#pragma unroll
    for (int j = 0; j < LD_TILE / B_NT; ++j) {
      const int p = t0 + warp * (LD_TILE / B_NW) + j * 32 + lane;
      const bool valid = p < t1;
      int x = v[j];
      int key = 0x10000 + lane;      // lanes without an in-window value: a key nobody shares
      if (valid) {
        if (do_gc && p < s20) {
          int g;
          if (r20 >= 2 && p >= q0 && p < q0 + r20) { x = rd[p + GC_WIN - r20]; g = gstar; }
          else { const int lo = gc_lo(p, L) - wlo; g = (int)pre[lo + GC_WIN] - (int)pre[lo]; }
          x = (int)(__dadd_rn(__ddiv_rn(__dmul_rn((double)x, mean), tab[g]), 0.5));
        }
        // value histogram over every position (N bases included)
        const int w = x - hb;
        if ((unsigned)w < (unsigned)B_K) key = w;
        else if (x == 0) key = 0x8000;                 // zeros below the window (N stretches): merged per warp, counted in a register
        else if (x >= HIST_ALL_BINS || x < 0) bad = 1;
        else atomicAdd(&hist_all[x], 1u);
        // N-compacted store
        int cidx;
        if (!hasn) cidx = p - shift0;
        else {
          int k = k0, sh = shift0;
          while (k < nn && nend[k] < p) { sh += nend[k] - nbeg[k] + 1; ++k; }
          cidx = (k < nn && p >= nbeg[k]) ? -1 : p - sh;
        }
        if (cidx >= 0) rdc[cidx] = x;
      }
      const unsigned m = __match_any_sync(0xffffffffu, key);
      if (lane == __ffs((int)m) - 1) { if (key < B_K) wh[key] += (u32)__popc(m); else if (key == 0x8000) zeros += (u32)__popc(m); }
      __syncwarp();
    }
  }
  c.sync();
  for (int w = tid; w < B_K; w += B_NT) {
    u32 s = 0;
    for (int k = 0; k < B_NW; ++k) s += vh[k * B_K + w];
    if (s) atomicAdd(&hist_all[hb + w], s);
  }
  bad = c.reduce(bad, MaxOp());
  zeros = c.reduce(zeros, SumOp());
  if (tid == 0 && zeros) atomicAdd(&hist_all[0], zeros);
  if (tid == 0 && bad) atomicOr(&st->err, (int)ERR_HIST_RANGE);
}

// ---------------------------------------------------------------------------------------------
// First histogram bucket whose running count reaches each of three ranks (the rule of
// partition_stat_tp, wufunctions.cpp:398-412: `run < r && run + cnt >= r`), over a value-indexed
// count array.  Results (bucket indices, -1 if never reached) to every thread.
template <class CT>
__device__ void cta_hist_pick(const Cta& c, const CT* hist, int nbins, u64 r0, u64 r1, u64 r2, int out[3], int* first_nz, int* last_nz) {
  const int chunk = (nbins + c.nthr - 1) / c.nthr;
  const int b0 = imin(c.tid * chunk, nbins), b1 = imin(b0 + chunk, nbins);
  i64 local = 0;
  int fnz = 0x7fffffff, lnz = -1;
  for (int b = b0; b < b1; ++b) { const u64 h = (u64)hist[b]; local += (i64)h; if (h) { if (fnz == 0x7fffffff) fnz = b; lnz = b; } }
  i64 tot;
  i64 run = c.scan_excl(local, &tot);
  int pick[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
  for (int b = b0; b < b1; ++b) {
    const u64 cnt = (u64)hist[b], before = (u64)run, after = before + cnt;
    if (before < r0 && after >= r0) pick[0] = b;
    if (before < r1 && after >= r1) pick[1] = b;
    if (before < r2 && after >= r2) pick[2] = b;
    run += (i64)cnt;
  }
  for (int k = 0; k < 3; ++k) { int v = c.reduce(pick[k], MinOp()); out[k] = v == 0x7fffffff ? -1 : v; }
  *first_nz = c.reduce(fnz, MinOp());
  *last_nz = c.reduce(lnz, MaxOp());
}

// apply_cap's parameters: median of the adjusted depths over all L positions, threshold med*cap,
// replacement value int(med*cap); also the value range of the class histograms of k_bins.
__global__ void __launch_bounds__(1024) k_cap_params(const u32* hist_all, DevState* st, int chist_rcap) {
  RSI_CTA_SETUP(c);
  const u64 n = (u64)st->L;
  int pick[3], fnz, lnz;
  cta_hist_pick(c, hist_all, (int)HIST_ALL_BINS, n / 4, n / 2, n * 3 / 4, pick, &fnz, &lnz);
  if (c.tid == 0) {
    double med = (fnz == lnz || pick[1] < 0) ? (double)fnz : (double)pick[1];  // all equal -> the mean, i.e. that value
    st->cap_median = med;
    int R = lnz + 1;
    if (st->cap_on) {
      st->cap_thr = med * st->cap;
      st->capv = (int)(med * st->cap);
      // values above the threshold become capv; the largest surviving value is <= floor(cap_thr)
      int top = (int)st->cap_thr; if (top < st->capv) top = st->capv;
      if (R > top + 1) R = top + 1;
    } else { st->cap_thr = 1e300; st->capv = 0; }
    if (R > chist_rcap) { st->err |= ERR_HIST_RANGE; R = chist_rcap; }
    st->chist_R = R;
  }
}

}  // namespace rsigpu
