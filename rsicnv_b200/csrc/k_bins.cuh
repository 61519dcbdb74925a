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

enum { C_NT = 512, C_NW = C_NT / 32, C_K = 128, C_KP = C_K + 2 /* padded row: 65 words, lanes of a step hit 31 different banks */, C_TP = 8192 /* max staged bases */ };
#define RSI_SMEM_C ((size_t)C_NW * MAD_CLASSES * C_KP * 2 + (size_t)C_TP * 4)

__device__ __forceinline__ i64 warp_sum_i64(i64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// chist[c * R + v]: #bases of class c (compacted index mod 31, index < 31*floor(Lc/31)) with capped
// value v; thist[v]: the < 31 bases beyond.  bins_per_tile * m <= C_TP.
__global__ void __launch_bounds__(C_NT) k_bins(int* __restrict__ rdc, float* __restrict__ bin_med, int* __restrict__ bin_medint,
                                                i64* __restrict__ bin_sum, u32* chist, u32* thist, DevState* st, int bins_per_tile) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  u16* ct = reinterpret_cast<u16*>(smem);                  // [C_NW][MAD_CLASSES][C_KP]
  int* vals = reinterpret_cast<int*>(smem + (size_t)C_NW * MAD_CLASSES * C_KP * 2);
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
  const int ntiles = nb > 0 ? (nb + bins_per_tile - 1) / bins_per_tile : 1;
  for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x) {
    const int b0 = tile * bins_per_tile, nbt = imin(bins_per_tile, nb - b0);
    const int B = b0 * m;
    const int np = (tile == ntiles - 1) ? Lc - B : nbt * m;   // the last tile also owns the tail beyond nb*m
    c.sync();
    for (int q0 = 0; q0 < np; q0 += C_TP) {                    // (only the tail of a tiny contig can exceed C_TP)
      const int nq = imin(C_TP, np - q0);
      if (q0) c.sync();
      for (int qb = 0; qb < nq; qb += C_NT * 4) {   // 4 independent loads in flight per thread before any (aliasing) store
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int q = qb + k * C_NT + tid; v[k] = q < nq ? rdc[B + q0 + q] : 0; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int q = qb + k * C_NT + tid;
          if (q >= nq) continue;
          if (cap_on && v[k] > ithr) { v[k] = capv; rdc[B + q0 + q] = capv; }
          vals[q] = v[k];
        }
      }
      c.sync();
      {
        const int Bq = B + q0;
        // warp-steps of 31 consecutive bases: 31 distinct classes, lane 31 idles
        for (int s0 = warp * 31; s0 < nq; s0 += C_NW * 31) {
          const int q = s0 + lane;
          if (lane < 31 && q < nq) {
            const int v = vals[q], w = v - wb, cls = (Bq + q) % MAD_CLASSES;
            if (Bq + q >= sub31) { if (v >= 0 && v < R) atomicAdd(&thist[v], 1u); }
            else if ((unsigned)w < (unsigned)C_K) wt[cls * C_KP + w] += 1;
            else if (v >= 0 && v < R) atomicAdd(&chist[cls * R + v], 1u);
          }
          __syncwarp();
        }
      }
      if (q0 == 0) {
        for (int b = warp; b < nbt; b += C_NW) {
          const int* x = vals + b * m;
          const int need = (m - 1) / 2 + 1;                    // smallest v with #{x <= v} >= need (m is odd)
          if (m <= 128) {   // the bin lives in 4 registers per lane; one full-mask REDUX per probe
            int r[4]; int lo = 0x7fffffff, hi = -0x7fffffff - 1, s32 = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int j = lane + 32 * k; r[k] = j < m ? x[j] : 0x7fffffff; if (j < m) { lo = imin(lo, r[k]); hi = imax(hi, r[k]); s32 += r[k]; } }
            lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
            const i64 s = (i64)__reduce_add_sync(0xffffffffu, (unsigned)s32);     // < 128 * 2^24
            while (lo < hi) {
              const int mid = lo + ((hi - lo) >> 1);
              const int cnt = __reduce_add_sync(0xffffffffu, (r[0] <= mid) + (r[1] <= mid) + (r[2] <= mid) + (r[3] <= mid));
              if (cnt >= need) hi = mid; else lo = mid + 1;
            }
            if (lane == 0) { bin_med[b0 + b] = (float)lo; bin_medint[b0 + b] = lo; bin_sum[b0 + b] = s; mx = lmax(mx, s); }
            continue;
          }
          int lo = 0x7fffffff, hi = -0x7fffffff - 1; i64 s = 0;
          for (int j = lane; j < m; j += 32) { const int v = x[j]; lo = imin(lo, v); hi = imax(hi, v); s += v; }
          lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
          s = warp_sum_i64(s);
          while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            int cnt = 0;
            for (int j0 = 0; j0 < m; j0 += 32) { const int j = j0 + lane; cnt += __popc(__ballot_sync(0xffffffffu, j < m && x[j] <= mid)); }
            if (cnt >= need) hi = mid; else lo = mid + 1;
          }
          if (lane == 0) { bin_med[b0 + b] = (float)lo; bin_medint[b0 + b] = lo; bin_sum[b0 + b] = s; mx = lmax(mx, s); }
        }
      }
    }
  }
  c.sync();
  for (int item = tid; item < MAD_CLASSES * C_K; item += C_NT) {
    const int cl = item / C_K, w = item % C_K;
    u32 s = 0;
    for (int k = 0; k < C_NW; ++k) s += ct[((size_t)k * MAD_CLASSES + cl) * C_KP + w];
    if (s && wb + w < R) atomicAdd(&chist[cl * R + wb + w], s);
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
