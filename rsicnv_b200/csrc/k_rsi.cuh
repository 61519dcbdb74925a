// k_rsi.cuh -- Robust Segment Identification on the bin arrays.
//
// Replaces (reference file:line relative to src/):
//   rsicnvnbn / rsicnvmed (scalars)   rsi.cpp:1262-1360, 1402-1515   k_rsi_params1, k_rsi_params2
//   rsistatus + runmeantp             rsi.cpp:1191-1259, wufunctions.cpp:572-647   k_rsi_scan, k_rsi_cnt_del, k_rsi_cnt_dup, k_rsi_status
//   filterstatus_tp                   rsi.cpp:948-1057               k_nz_scatter, k_level0_chain*, k_level_sums, k_filter_params, k_filter_trim
//   get_continuous_segments           rsi.cpp:291-326                run detection inside k_rsi_status / k_filter_trim / k_runs_scatter
//   get_rsi_segments                  rsi.cpp:1060-1117              k_runs_scatter, k_run_argmax
//
// rsistatus is evaluated in its order-free form: every (L, i) window is tested
// independently on an EXACT window sum (fixed-point prefix sums in shared memory), hits record the
// smallest covering L per bin with shared/global atomicMin, and the reference's "L ascending,
// first writer wins, stop once 20% is marked" rule is applied afterwards from a histogram over L.
#pragma once
#include "k_quant.cuh"

namespace rsigpu {

enum { S_NT = 256, S_T = 1024, S_H = LMAX_CAP / 2 + 2 };
enum { MINL_INF = 0x7f7f7f7f };


// ---- scalars before the first rsistatus pass (which: 0 = NBN on the transformed bins, 1 = MED on the medians)
__global__ void k_rsi_params1(const float* __restrict__ t, DevState* st, int which, double threshold) {
  if (threadIdx.x || blockIdx.x) return;
  const double tmedian = which == 0 ? st->qj[0].q[1] : st->rdmedian;
  st->tmedian = tmedian;
}
__global__ void k_rsi_params2(const float* __restrict__ t, DevState* st, int which, double threshold, int slot) {
  if (threadIdx.x || blockIdx.x) return;
  const double tmedian = st->tmedian;
  const double tsigma = st->qj[slot].q[1] / 0.6745;
  double tlamda = st->factor * tsigma, target, dev;
  int calmax;
  if (which == 0) {
    const float d20 = t[2] - t[0];
    target = (double)d20 * sqrt(2.5);
    tlamda = tlamda > target ? tlamda : target;
    const double dnb = (double)fabsf(d20) + 0.0001;
    const double x = tlamda * 2 / dnb;
    calmax = (int)(x * x);
    dev = tsigma * 3.0;
  } else {
    target = tmedian * sqrt(2.0);
    tlamda = tlamda > target ? tlamda : target;
    if (threshold > 0) tlamda = tmedian * threshold;
    const double x = tlamda * 4 / (tmedian + 0.001);
    calmax = (int)(x * x);
    dev = tmedian * 0.6;
  }
  int Lmax = st->Lmax_base;
  if (Lmax < calmax) Lmax = calmax;
  if (Lmax > LMAX_CAP) { st->err |= ERR_LMAX; Lmax = LMAX_CAP; }
  st->tsigma = tsigma; st->tlamda = tlamda; st->target = target; st->dev = dev; st->Lmax = Lmax;
  { RsiLogT& g = st->rlog[which]; g.tmedian1 = tmedian; g.tsigma1 = tsigma; g.tlamda1 = tlamda; g.target = target; g.calmax = calmax; g.lmax = Lmax; g.used = 1; g.filt_on = -1; }
  for (int L = 0; L < LMAX_CAP + 2; ++L) { st->cnt_del[L] = 0; st->cnt_dup[L] = 0; }
  for (int l = 0; l < 2 * LMAX_CAP + 3; ++l) { st->lvl_sum[l] = 0.f; st->lvl_cnt[l] = 0; }
  st->st_lo = 0; st->st_hi = 0; st->last_run_start = -1; st->n_nonzero = 0;
}
// re-estimation on the unmarked bins after filterstatus (rsi.cpp:1307-1318 / 1457-1468)
__global__ void k_rsi_params3(DevState* st, int slot_med, int slot_sig, int which) {
  if (threadIdx.x || blockIdx.x) return;
  const u32 k = st->qj[slot_med].n;
  st->n_unmarked = k;
  if (k > (u32)(st->nb / 2)) {
    st->tmedian = st->qj[slot_med].q[1];
    st->tsigma = st->qj[slot_sig].q[1] / 0.6745;
    const double tl = st->factor * st->tsigma;
    st->tlamda = tl > st->target ? tl : st->target;
  }
  st->out_tmedian = st->tmedian; st->out_tlamda = st->tlamda;
  { RsiLogT& g = st->rlog[which]; g.tmedian2 = st->tmedian; g.tsigma2 = st->tsigma; g.tlamda2 = st->tlamda; }
  for (int L = 0; L < LMAX_CAP + 2; ++L) { st->cnt_del[L] = 0; st->cnt_dup[L] = 0; }
  st->st_lo = 0; st->st_hi = 0; st->last_run_start = -1; st->n_nonzero = 0;
}

// ---------------------------------------------------------------------------------------------
// window-median test of rsistatus (alglib::median over RDmedint[i1..i2], rsi.cpp:1209/1238) from
// prefix counts; only the rare tie case (exactly half of an even window on each side of the limit)
// looks at the values.
__device__ bool window_median_ok(const int* __restrict__ medint, int i1, int L, int c_in, double lim, int sign) {
  // sign < 0 (DEL): c_in = #{v <= lim}, pass iff median <= lim;  sign > 0 (DUP): c_in = #{v >= lim}, pass iff median >= lim
  if (L & 1) return c_in >= (L - 1) / 2 + 1;
  if (c_in >= L / 2 + 1) return true;
  if (c_in <= L / 2 - 1) return false;
  int a, b;
  if (sign < 0) {
    a = -0x7fffffff - 1; b = 0x7fffffff;  // a = max{v <= lim}, b = min{v > lim}
    for (int j = 0; j < L; ++j) { const int v = medint[i1 + j]; if ((double)v <= lim) a = imax(a, v); else b = imin(b, v); }
    return 0.5 * ((double)a + (double)b) <= lim;
  }
  a = -0x7fffffff - 1; b = 0x7fffffff;    // a = max{v < lim}, b = min{v >= lim}
  for (int j = 0; j < L; ++j) { const int v = medint[i1 + j]; if ((double)v >= lim) b = imin(b, v); else a = imax(a, v); }
  return 0.5 * ((double)a + (double)b) >= lim;
}

// One block OWNS S_T bins and evaluates every window (centre in the tile or within Lmax/2 of it, every
// length) that can mark them.  The reference's end trimming (rsi.cpp:1212-1215 / 1241-1244) moves a
// window [s, e] to [nextB(nextA(s)), prevB(prevA(e))] (A: bin value on the far side of tmedian, B: bin
// median beyond the limit), so bin j is written by that window iff
//        s <= SA(j) = prevA(prevB(j))   and   e >= EA(j) = nextA(nextB(j)),
// i.e. iff SOME hit centre lies in [EA(j) - (L-1) + L/2, SA(j) + L/2].  Hits of one length are kept as a
// bit per centre (one __ballot_sync word per 32 centres), each thread tests the centre range of the
// bins it owns and remembers the first (smallest) L in registers: no atomics, no per-hit loops, and the
// result does not depend on how hits cluster.  Dynamic shared memory (NB = S_T + 2*LMAX_CAP staged bins):
//   P[NB+1] i64 | SA/EA x DEL/DUP [4][S_T] i32 | cle,cge [NB+1] u16 | hit words [2 buffers][2 lengths][2 signs][S_NC/32+1] u32
enum { S_NB = S_T + 2 * LMAX_CAP, S_NC = S_T + 2 * S_H };
enum { LMAX_SMALL = 256 };        // most contigs need far fewer window lengths than LMAX_CAP: a second instantiation with a small footprint
#define RSI_SCAN_SMEM_T(LCAP) ((size_t)(S_T + 2 * (LCAP) + 1) * 8 + (size_t)4 * S_T * 4 + (size_t)2 * (S_T + 2 * (LCAP) + 2) * 2 + (size_t)8 * ((S_T + 2 * ((LCAP) / 2 + 2)) / 32 + 2) * 4)
#define RSI_SCAN_SMEM RSI_SCAN_SMEM_T(LMAX_CAP)

// prev / next index with a condition, over the staged bins (block-wide max / min scans, thread-contiguous)
__device__ void scan_prev_next(const Cta& c, const u8* cond, int N, int* prevv, int* nextv) {
  const int per = (N + c.nthr - 1) / c.nthr;
  const int k0 = imin(c.tid * per, N), k1 = imin(k0 + per, N);
  int last = -1;
  for (int k = k0; k < k1; ++k) if (cond[k]) last = k;
  // exclusive max-scan of `last` across threads
  int v = last;
  const int lane = c.tid & 31, warp = c.tid >> 5, nw = (c.nthr + 31) >> 5;
  for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = imax(v, u); }
  int* slots = reinterpret_cast<int*>(c.red);
  c.sync(); if (lane == 31) slots[warp] = v; c.sync();
  int pre = -1; for (int w = 0; w < warp; ++w) pre = imax(pre, slots[w]);
  int up = __shfl_up_sync(0xffffffffu, v, 1);
  int run = imax(pre, lane ? up : -1);
  for (int k = k0; k < k1; ++k) { if (cond[k]) run = k; prevv[k] = run; }
  c.sync();
  // reverse: first index >= k with cond
  int first = 0x7fffffff;
  for (int k = k1 - 1; k >= k0; --k) if (cond[k]) first = k;
  v = first;
  for (int o = 1; o < 32; o <<= 1) { int u = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v = imin(v, u); }
  if (lane == 0) slots[warp] = v; c.sync();
  int post = 0x7fffffff; for (int w = warp + 1; w < nw; ++w) post = imin(post, slots[w]);
  int dn = __shfl_down_sync(0xffffffffu, v, 1);
  run = imin(post, lane < 31 ? dn : 0x7fffffff);
  for (int k = k1 - 1; k >= k0; --k) { if (cond[k]) run = k; nextv[k] = run; }
  c.sync();
}

// The hit tests of rsistatus, (float(sum/L) - tmedian)*sqrt(L) <= -tlamda (DEL) / >= tlamda (DUP), are monotone
// in the exact window sum (every rounding step is monotone), so for each window length they are equivalent
// to integer thresholds on the fixed-point sum: thr[2L] = largest sum that is a DEL hit (-1: none),
// thr[2L+1] = smallest sum that is a DUP hit (INT64_MAX: none).  One thread per length, binary search with the
// reference's own arithmetic.
__device__ __forceinline__ double rsi_score(i64 s, double dL, double sL, double tmed) {
  const float mean = (float)__ddiv_rn((double)s * (1.0 / 68719476736.0), dL);
  return __dmul_rn((double)mean - tmed, sL);
}
__global__ void k_rsi_thresholds(i64* __restrict__ thr, DevState* st) {
  const int Lmax = st->Lmax;
  const double tmed = st->tmedian, tlam = st->tlamda;
  for (int L = 1 + (int)(blockIdx.x * blockDim.x + threadIdx.x); L <= Lmax; L += (int)(gridDim.x * blockDim.x)) {
    const double dL = (double)L, sL = sqrt(dL);
    const i64 smax = (i64)L * 16384ll * 68719476736ll;      // bin values are < 2^14 (checked by the scan kernel)
    // DEL: hits form a prefix [0, S]
    i64 lo = -1, hi = smax;                                   // invariant: lo is a hit (or -1), hi + 1 is not
    if (!(rsi_score(smax, dL, sL, tmed) > -tlam)) lo = smax;
    else while (lo < hi) { const i64 mid = lo + (hi - lo + 1) / 2; if (!(rsi_score(mid, dL, sL, tmed) > -tlam)) lo = mid; else hi = mid - 1; }
    thr[2 * L] = lo;
    // DUP: hits form a suffix [S, smax]
    i64 a = 0, b = smax + 1;                                  // smallest hit in [a, b]; b = smax + 1 means none
    while (a < b) { const i64 mid = a + (b - a) / 2; if (!(rsi_score(mid, dL, sL, tmed) < tlam)) b = mid; else a = mid + 1; }
    thr[2 * L + 1] = a > smax ? 0x7fffffffffffffffll : a;
  }
}

template <int LCAP, bool SMALL>
__device__ __forceinline__ void rsi_scan_body(const float* __restrict__ t, const int* __restrict__ medint, u32* __restrict__ minl_del,
                                              u32* __restrict__ minl_dup, int* __restrict__ scratch, const i64* __restrict__ thr, DevState* st) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  constexpr int S_NBT = S_T + 2 * LCAP, S_NCT = S_T + 2 * (LCAP / 2 + 2);
  i64* P = reinterpret_cast<i64*>(smem);
  int* SAEA = reinterpret_cast<int*>(P + (S_NBT + 1));        // [0]=SA_del [1]=EA_del [2]=SA_dup [3]=EA_dup, S_T each
  u16* cle = reinterpret_cast<u16*>(SAEA + 4 * S_T);
  u16* cge = cle + (S_NBT + 2);
  u32* hw = reinterpret_cast<u32*>(cge + (S_NBT + 2));              // [buf][sign][word]
  const int NW = S_NCT / 32 + 2;
  const int nb = st->nb, Lmax = st->Lmax;
  if (SMALL ? Lmax > LCAP : Lmax <= LMAX_SMALL) return;             // the other instantiation handles this contig
  const double tmed = st->tmedian, limd = st->lim_del, limu = st->lim_dup;
  const int c0 = (int)blockIdx.x * S_T;
  const int HB = Lmax + 1;                        // staged bins: [c0 - HB, c0 + S_T + HB)
  const int base = c0 - HB;                       // smem index k <-> bin base + k
  const int N = S_T + 2 * HB;
  const int HC = Lmax / 2 + 2;                    // centres:     [c0 - HC, c0 + S_T + HC)
  const int cbase = c0 - HC;
  const int NC = S_T + 2 * HC;
  const int tid = c.tid, lane = tid & 31, warp = tid >> 5;
  // per-block global scratch for the prev/next arrays (8 x N ints) and condition bytes
  int* sc = scratch + (size_t)blockIdx.x * (8 * (size_t)S_NB + 2 * (size_t)S_NB);
  int* pv[4]; int* nx[4];
  for (int k = 0; k < 4; ++k) { pv[k] = sc + (size_t)(2 * k) * S_NB; nx[k] = sc + (size_t)(2 * k + 1) * S_NB; }
  u8* cond = reinterpret_cast<u8*>(sc + 8 * (size_t)S_NB);   // 4 x S_NB bytes: A_del, B_del, A_dup, B_dup
  // ---- stage: fixed-point values, limit flags, block-wide exclusive prefix, trimming conditions
  const int chunk = (N + S_NT - 1) / S_NT;
  const int k0 = imin(tid * chunk, N), k1 = imin(k0 + chunk, N);
  i64 ls = 0; int ld = 0, lu = 0, bad = 0; float tmax = 0.f; int lowexp = 1000;
  for (int k = k0; k < k1; ++k) {
    const int b = base + k;
    u8 ca = 0, cb = 0, cc = 0, cd = 0;
    if (b >= 0 && b < nb) {
      const float x = t[b];
      const i64 fx = (i64)((double)x * 68719476736.0);    // * 2^FX_SHIFT
      if ((double)fx * (1.0 / 68719476736.0) != (double)x || !(x >= 0.f) || x >= 16384.f) bad = 1;
      if (x > 0.f) {
        const u32 u = __float_as_uint(x);
        const int e = (int)((u >> 23) & 0xff) - 127 - 23;
        const u32 man = (u & 0x7fffff) | 0x800000;
        lowexp = imin(lowexp, e + (__ffs((int)man) - 1));
        tmax = x > tmax ? x : tmax;
      }
      ls += fx;
      const int v = medint[b];
      ld += ((double)v <= limd) ? 1 : 0; lu += ((double)v >= limu) ? 1 : 0;
      ca = !((double)x > tmed); cb = !((double)v > limd);      // DEL loops run while t > tmed / medint > lim
      cc = !((double)x < tmed); cd = !((double)v < limu);      // DUP loops run while t < tmed / medint < lim
    }
    cond[k] = ca; cond[S_NB + k] = cb; cond[2 * S_NB + k] = cc; cond[3 * S_NB + k] = cd;
  }
  i64 tots; int totd, totu;
  i64 rs = c.scan_excl(ls, &tots);
  int rd_ = c.scan_excl(ld, &totd);
  int ru = c.scan_excl(lu, &totu);
  for (int k = k0; k < k1; ++k) {
    const int b = base + k;
    P[k] = rs; cle[k] = (u16)rd_; cge[k] = (u16)ru;
    if (b >= 0 && b < nb) {
      rs += (i64)((double)t[b] * 68719476736.0);
      const int v = medint[b];
      rd_ += ((double)v <= limd) ? 1 : 0; ru += ((double)v >= limu) ? 1 : 0;
    }
  }
  if (tid == 0) { P[N] = tots; cle[N] = (u16)totd; cge[N] = (u16)totu; }
  // exactness of the reference's own double window sums: all partial sums need <= 53 significant bits
  tmax = c.reduce(tmax, MaxOp()); lowexp = c.reduce(lowexp, MinOp()); bad = c.reduce(bad, MaxOp());
  if (tid == 0) {
    const int span = Lmax > 8192 ? Lmax : 8192;
    if (bad || (lowexp < 1000 && (double)tmax * (double)span >= ldexp(1.0, 53 + lowexp))) atomicOr(&st->err, (int)ERR_FIXEDPOINT);
  }
  c.sync();
  for (int k = 0; k < 4; ++k) scan_prev_next(c, cond + (size_t)k * S_NB, N, pv[k], nx[k]);
  // SA(j) = prevA(prevB(j)), EA(j) = nextA(nextB(j)) for the owned bins (smem indices; -1 / INF = none in range)
  for (int r = tid; r < S_T; r += S_NT) {
    const int k = HB + r;
    for (int sgn = 0; sgn < 2; ++sgn) {
      const int pb = pv[2 * sgn + 1][k];
      SAEA[(2 * sgn) * S_T + r] = pb < 0 ? -1 : pv[2 * sgn][pb];
      const int nbk = nx[2 * sgn + 1][k];
      SAEA[(2 * sgn + 1) * S_T + r] = nbk == 0x7fffffff ? 0x7fffffff : nx[2 * sgn][nbk];
    }
  }
  c.sync();
  u32 ml_del[S_T / S_NT], ml_dup[S_T / S_NT];
#pragma unroll
  for (int r = 0; r < S_T / S_NT; ++r) { ml_del[r] = MINL_INF; ml_dup[r] = MINL_INF; }
  const int nwords = (NC + 31) / 32;
  // ---- every window length.  Lengths 2k and 2k+1 have the same half-width k (same centres, same window start), so they are
  // taken together: three prefix loads give both sums, and there is one barrier per PAIR.  L = 1 goes alone.
  int pbuf = 0;
  for (int L0 = 1; L0 <= Lmax; L0 += (L0 == 1 ? 1 : 2)) {
    const int nL = (L0 == 1 || L0 + 1 > Lmax) ? 1 : 2;
    const int h = L0 / 2;
    const i64 sdel0 = thr[2 * L0], sdup0 = thr[2 * L0 + 1];
    const i64 sdel1 = nL == 2 ? thr[2 * L0 + 2] : -1, sdup1 = nL == 2 ? thr[2 * L0 + 3] : 0x7fffffffffffffffll;
    const int ilo = h + 1, ihi = nb - h - 1;        // valid centres: ilo <= i < ihi
    u32* hwb = hw + (size_t)(pbuf * 4) * NW;        // double-buffered hit words [length][sign][word]: one barrier per pair
    pbuf ^= 1;
    // "sum <= sdel or sum >= sdup" as ONE unsigned compare: (sum - sdel - 1) mod 2^64 >= sdup - sdel - 1 (a range of 0 = always look closer)
    const u64 rng0 = sdel0 < sdup0 ? (u64)sdup0 - (u64)sdel0 - 1ull : 0ull, rng1 = sdel1 < sdup1 ? (u64)sdup1 - (u64)sdel1 - 1ull : 0ull;
    int anyhit = 0;
    for (int w = warp; w < nwords; w += S_NT / 32) {
      const int ci = w * 32 + lane;                 // centre index within [0, NC)
      const int i = cbase + ci;
      bool hd0 = false, hu0 = false, hd1 = false, hu1 = false, near = false;
      const bool valid = ci < NC && i >= ilo && i < ihi;
      const int kw = i - h - base;                  // smem index of the window start
      i64 s0 = 0, s1 = 0;
      if (valid) {
        const i64 p0 = P[kw];
        s0 = P[kw + L0] - p0;
        near = (u64)s0 - (u64)sdel0 - 1ull >= rng0;
        if (nL == 2) { s1 = P[kw + L0 + 1] - p0; near = near || ((u64)s1 - (u64)sdel1 - 1ull >= rng1); }
      }
      if (!__any_sync(0xffffffffu, near)) {         // no window of this word reaches a threshold (nearly always)
        if (lane == 0) { hwb[w] = 0u; hwb[NW + w] = 0u; hwb[2 * NW + w] = 0u; hwb[3 * NW + w] = 0u; }
        continue;
      }
      if (valid) {
        hd0 = s0 <= sdel0; hu0 = s0 >= sdup0;
        if (hd0) hd0 = window_median_ok(medint, i - h, L0, (int)cle[kw + L0] - (int)cle[kw], limd, -1);
        if (hu0) hu0 = window_median_ok(medint, i - h, L0, (int)cge[kw + L0] - (int)cge[kw], limu, +1);
        if (nL == 2) {
          hd1 = s1 <= sdel1; hu1 = s1 >= sdup1;
          if (hd1) hd1 = window_median_ok(medint, i - h, L0 + 1, (int)cle[kw + L0 + 1] - (int)cle[kw], limd, -1);
          if (hu1) hu1 = window_median_ok(medint, i - h, L0 + 1, (int)cge[kw + L0 + 1] - (int)cge[kw], limu, +1);
        }
      }
      const u32 bd0 = __ballot_sync(0xffffffffu, hd0), bu0 = __ballot_sync(0xffffffffu, hu0);
      const u32 bd1 = __ballot_sync(0xffffffffu, hd1), bu1 = __ballot_sync(0xffffffffu, hu1);
      if (lane == 0) { hwb[w] = bd0; hwb[NW + w] = bu0; hwb[2 * NW + w] = bd1; hwb[3 * NW + w] = bu1; }
      anyhit |= (bd0 | bu0 | bd1 | bu1) != 0u;
    }
    if (!__syncthreads_or(anyhit)) continue;        // no window of these lengths hits anywhere near the tile (the common case)
    for (int q = 0; q < nL; ++q) {                  // increasing length: the smallest L wins
      const int L = L0 + q;
      const u32* hwd = hwb + (size_t)(2 * q) * NW;
      const u32* hwu = hwd + NW;
#pragma unroll
      for (int r = 0; r < S_T / S_NT; ++r) {
        const int jr = tid + r * S_NT;                // owned bin, smem index HB + jr
        const int j = c0 + jr;
        if (j >= nb) continue;
#pragma unroll
        for (int sgn = 0; sgn < 2; ++sgn) {
          if ((sgn ? ml_dup[r] : ml_del[r]) != MINL_INF) continue;
          const int sa = SAEA[(2 * sgn) * S_T + jr], ea = SAEA[(2 * sgn + 1) * S_T + jr];
          if (sa < 0 || ea == 0x7fffffff) continue;
          // windows [s, e] (smem indices) with s <= sa, e = s + L - 1 >= ea; centre smem index = s + h
          int slo = ea - (L - 1), shi = sa;
          if (slo > shi) continue;
          // to centre indices within [0, NC): centre bin = base + s + h  =>  ci = s + h + base - cbase
          int clo = slo + h + (base - cbase), chi = shi + h + (base - cbase);
          if (clo < 0) clo = 0;
          if (chi > NC - 1) chi = NC - 1;
          if (clo > chi) continue;
          const u32* hwp = sgn ? hwu : hwd;
          bool any = false;
          for (int w = clo >> 5; w <= (chi >> 5) && !any; ++w) {
            u32 m = hwp[w];
            if (w == (clo >> 5)) m &= 0xffffffffu << (clo & 31);
            if (w == (chi >> 5)) m &= 0xffffffffu >> (31 - (chi & 31));
            any = m != 0;
          }
          if (any) { if (sgn) ml_dup[r] = (u32)L; else ml_del[r] = (u32)L; }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < S_T / S_NT; ++r) {
    const int j = c0 + tid + r * S_NT;
    if (j < nb) { minl_del[j] = ml_del[r]; minl_dup[j] = ml_dup[r]; }
  }
}

__global__ void __launch_bounds__(S_NT) k_rsi_scan(const float* __restrict__ t, const int* __restrict__ medint, u32* __restrict__ minl_del,
                                                    u32* __restrict__ minl_dup, int* __restrict__ scratch, const i64* __restrict__ thr, DevState* st) {
  rsi_scan_body<LMAX_CAP, false>(t, medint, minl_del, minl_dup, scratch, thr, st);
}
__global__ void __launch_bounds__(S_NT, 4) k_rsi_scan_small(const float* __restrict__ t, const int* __restrict__ medint, u32* __restrict__ minl_del,
                                                          u32* __restrict__ minl_dup, int* __restrict__ scratch, const i64* __restrict__ thr, DevState* st) {
  rsi_scan_body<LMAX_SMALL, true>(t, medint, minl_del, minl_dup, scratch, thr, st);
}

// ---- the "stop once more than 20% is marked" rule (rsi.cpp:1225, 1255) from the histogram over L
__device__ int rsi_lbreak(const u32* cnt, int Lmax, int nb) {
  u64 cum = 0;
  for (int L = 1; L <= Lmax; ++L) { cum += cnt[L]; if ((double)cum / (double)nb > 0.2) return L; }
  return Lmax;
}
__global__ void k_rsi_cnt_del(const u32* __restrict__ minl_del, DevState* st) {
  const int nb = st->nb, Lmax = st->Lmax;
  for (int j = (int)(blockIdx.x * blockDim.x + threadIdx.x); j < nb; j += (int)(gridDim.x * blockDim.x)) {
    const u32 l = minl_del[j];
    if (l <= (u32)Lmax) atomicAdd(&st->cnt_del[l], 1u);
  }
}
__global__ void k_rsi_cnt_dup(const u32* __restrict__ minl_del, const u32* __restrict__ minl_dup, DevState* st) {
  __shared__ int s_lb;
  const int nb = st->nb, Lmax = st->Lmax;
  if (threadIdx.x == 0) { s_lb = rsi_lbreak(st->cnt_del, Lmax, nb); if (blockIdx.x == 0) st->lbreak_del = s_lb; }
  __syncthreads();
  const u32 lb = (u32)s_lb;
  for (int j = (int)(blockIdx.x * blockDim.x + threadIdx.x); j < nb; j += (int)(gridDim.x * blockDim.x)) {
    if (minl_del[j] <= lb) continue;                 // written by the DEL pass: not writable any more
    const u32 l = minl_dup[j];
    if (l <= (u32)Lmax) atomicAdd(&st->cnt_dup[l], 1u);
  }
}
__device__ __forceinline__ int rsi_status_of(const u32* __restrict__ minl_del, const u32* __restrict__ minl_dup, int j, u32 lbd, u32 lbu) {
  const u32 a = minl_del[j];
  if (a <= lbd) return -(int)a;
  const u32 b = minl_dup[j];
  if (b <= lbu) return (int)b;
  return 0;
}
__device__ __forceinline__ bool same_run(int prev, int cur) { return prev != 0 && cur != 0 && ((prev > 0) == (cur > 0)); }

// status array + its range, the number of marked bins per 1024-bin tile (for the ordered
// compactions that follow) and the start of the last run.
__global__ void k_rsi_status(const u32* __restrict__ minl_del, const u32* __restrict__ minl_dup, int* __restrict__ status, int* __restrict__ tile_nz,
                             DevState* st) {
  RSI_CTA_SETUP(c);
  __shared__ int s_lb;
  const int nb = st->nb, Lmax = st->Lmax;
  if (c.tid == 0) { s_lb = rsi_lbreak(st->cnt_dup, Lmax, nb); if (blockIdx.x == 0) st->lbreak_dup = s_lb; }
  c.sync();
  const u32 lbd = (u32)st->lbreak_del, lbu = (u32)s_lb;
  int lo = 0, hi = 0, last = -1;
  const int ntiles = (nb + 1023) / 1024;
  for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x) {
    int nz = 0;
    for (int j = tile * 1024 + c.tid; j < imin(nb, tile * 1024 + 1024); j += c.nthr) {
      const int s = rsi_status_of(minl_del, minl_dup, j, lbd, lbu);
      status[j] = s;
      if (s) {
        ++nz; lo = imin(lo, s); hi = imax(hi, s);
        const int prev = j > 0 ? rsi_status_of(minl_del, minl_dup, j - 1, lbd, lbu) : 0;
        if (!same_run(prev, s)) last = imax(last, j);
      }
    }
    nz = c.reduce(nz, SumOp());
    if (c.tid == 0) tile_nz[tile] = nz;
  }
  lo = c.reduce(lo, MinOp()); hi = c.reduce(hi, MaxOp()); last = c.reduce(last, MaxOp());
  if (c.tid == 0) { atomicMin(&st->st_lo, lo); atomicMax(&st->st_hi, hi); atomicMax(&st->last_run_start, last); }
}

// ---------------------------------------------------------------------------------------------
// filterstatus_tp.  Ordered list of the marked bins (index order) for the per-level sums.
__global__ void k_nz_scatter(const float* __restrict__ t, const int* __restrict__ status, const int* __restrict__ tile_nz, int* __restrict__ nz_lvl,
                             float* __restrict__ nz_val, DevState* st) {
  RSI_CTA_SETUP(c);
  const int nb = st->nb;
  const int ntiles = (nb + 1023) / 1024;
  for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x) {
    int off = 0;
    for (int k = c.tid; k < tile; k += c.nthr) off += tile_nz[k];
    off = c.reduce(off, SumOp());
    if (tile == ntiles - 1 && c.tid == 0) st->n_nonzero = off + tile_nz[tile];
    if (tile_nz[tile] == 0) continue;
    for (int j0 = tile * 1024; j0 < imin(nb, tile * 1024 + 1024); j0 += c.nthr) {
      const int j = j0 + c.tid;
      const int f = (j < nb && j < tile * 1024 + 1024 && status[j] != 0) ? 1 : 0;
      int tot;
      const int ex = c.scan_excl(f, &tot);
      if (f) { nz_lvl[off + ex] = status[j]; nz_val[off + ex] = t[j]; }
      off += tot;
    }
  }
}

// Level-0 sum: the reference adds the unmarked bins one after the other into a FLOAT (rsi.cpp:967-974),
// so the result depends on the order and on every intermediate rounding.  Plain sequential form:
__global__ void k_level0_chain_seq(const float* __restrict__ t, const int* __restrict__ status, DevState* st) {
  if (threadIdx.x || blockIdx.x) return;
  const int nb = st->nb;
  float s = 0.f; u32 n = 0;
  for (int i = 0; i < nb; ++i) if (status[i] == 0) { s = __fadd_rn(s, t[i]); ++n; }
  const int l0 = -st->st_lo;
  st->lvl_sum[l0] = s; st->lvl_cnt[l0] = n;
}

// The same sum, exactly, with one thread block: while the accumulator S stays inside one binade its
// ulp u is fixed, S = N*u, and adding x = q*u + r moves N by q + [r > u/2] (+ the parity of N + q on
// an exact tie, round-to-nearest-even).  Such steps are functions of N's parity and compose
// associatively, so a block-wide scan applies thousands of additions at once; the scan stops at the
// element that carries S into the next binade, which is added with a real FADD, and restarts there.
struct ParStep { i64 d0, d1; };   // increment of N for incoming parity 0 / 1
__device__ __forceinline__ ParStep par_compose(ParStep a, ParStep b) {  // a first, then b
  ParStep r;
  r.d0 = a.d0 + ((a.d0 & 1) ? b.d1 : b.d0);
  r.d1 = a.d1 + (((1 + a.d1) & 1) ? b.d1 : b.d0);
  return r;
}
__device__ __forceinline__ ParStep par_elem(double x, double u, double inv_u) {
  // x >= 0; q = floor(x/u) clamped; r = x - q*u
  double qd = floor(x * inv_u);
  if (qd > 33554432.0) qd = 33554432.0;      // 2^25: anything >= 2^24 ends the binade anyway
  const i64 q = (i64)qd;
  const double r = x - qd * u;               // exact: u is a power of two and x has <= 24 significant bits
  const double half = 0.5 * u;
  ParStep s;
  if (r > half) { s.d0 = q + 1; s.d1 = q + 1; }
  else if (r == half) { s.d0 = q + (q & 1); s.d1 = q + ((q + 1) & 1); }
  else { s.d0 = q; s.d1 = q; }
  return s;
}
enum { CH_NT = 1024, CH_K = 8 };
__global__ void __launch_bounds__(CH_NT) k_level0_chain_scan(const float* __restrict__ t, const int* __restrict__ status, DevState* st) {
  RSI_CTA_SETUP(c);
  __shared__ ParStep s_w[CH_NT / 32];
  __shared__ float s_S; __shared__ int s_next;
  const int nb = st->nb, tid = c.tid, lane = tid & 31, warp = tid >> 5;
  float S = 0.f; int i0 = 0;
  u32 n = 0;
  for (int i = tid; i < nb; i += CH_NT) n += status[i] == 0 ? 1u : 0u;
  n = c.reduce(n, SumOp());
  while (i0 < nb) {
    if (!(S > 0.f)) {  // accumulator still zero (or not a positive normal): plain sequential adds until it is
      if (tid == 0) {
        int i = i0; float s = S;
        while (i < nb && !(s > 0.f)) { if (status[i] == 0) s = __fadd_rn(s, t[i]); ++i; }
        s_S = s; s_next = i;
      }
      c.sync(); S = s_S; i0 = s_next; c.sync();
      continue;
    }
    const int e = (int)((__float_as_uint(S) >> 23) & 0xff) - 127;   // S in [2^e, 2^(e+1))
    const double u = ldexp(1.0, e - 23), inv_u = ldexp(1.0, 23 - e);
    const i64 N0 = (i64)((double)S * inv_u);
    const int b0 = i0 + tid * CH_K;
    float xv[CH_K]; bool use[CH_K];
    ParStep acc; acc.d0 = 0; acc.d1 = 0;
#pragma unroll
    for (int k = 0; k < CH_K; ++k) {
      const int i = b0 + k;
      use[k] = i < nb && status[i] == 0;
      xv[k] = use[k] ? t[i] : 0.f;
      if (use[k]) acc = par_compose(acc, par_elem((double)xv[k], u, inv_u));
    }
    // inclusive scan of the per-thread steps across the block
    ParStep inc = acc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      ParStep up; up.d0 = __shfl_up_sync(0xffffffffu, inc.d0, o); up.d1 = __shfl_up_sync(0xffffffffu, inc.d1, o);
      if (lane >= o) inc = par_compose(up, inc);
    }
    c.sync();
    if (lane == 31) s_w[warp] = inc;
    c.sync();
    ParStep pre; pre.d0 = 0; pre.d1 = 0;
    for (int w = 0; w < warp; ++w) pre = par_compose(pre, s_w[w]);
    ParStep tot = pre;
    for (int w = warp; w < CH_NT / 32; ++w) tot = par_compose(tot, s_w[w]);
    // exclusive prefix of this thread = pre o (inclusive of the previous lane)
    ParStep prev; prev.d0 = __shfl_up_sync(0xffffffffu, inc.d0, 1); prev.d1 = __shfl_up_sync(0xffffffffu, inc.d1, 1);
    ParStep ex = lane == 0 ? pre : par_compose(pre, prev);
    const int p0 = (int)(N0 & 1);
    i64 N = N0 + (p0 ? ex.d1 : ex.d0);
    // walk the own elements with the actual N; the first element that leaves the binade ends the scan
    int cross = 0x7fffffff; float Sx = 0.f;
    if (N < 16777216) {
#pragma unroll
      for (int k = 0; k < CH_K; ++k) {
        if (!use[k] || cross != 0x7fffffff) continue;
        const ParStep s = par_elem((double)xv[k], u, inv_u);
        const i64 d = (N & 1) ? s.d1 : s.d0;
        if (N + d >= 16777216) { cross = b0 + k; Sx = __fadd_rn((float)((double)N * u), xv[k]); }
        else N += d;
      }
    }
    const int first = c.reduce(cross, MinOp());
    if (first == 0x7fffffff) {
      S = (float)((double)(N0 + (p0 ? tot.d1 : tot.d0)) * u);
      i0 += CH_NT * CH_K;
    } else {
      if (cross == first) s_S = Sx;
      c.sync();
      S = s_S; i0 = first + 1;
      c.sync();
    }
  }
  if (tid == 0) { const int l0 = -st->st_lo; st->lvl_sum[l0] = S; st->lvl_cnt[l0] = n; }
}

// Multi-block form of the same exact sum.  The array is cut into CHN_N chunks (one warp each).  A first
// pass gives every chunk an APPROXIMATE running total (fp64), hence a guess e0 of the accumulator's binade
// when the chain reaches it; the second pass composes the chunk's parity-step functions under the two
// hypotheses "binade e0" and "binade e0+1".  A final one-warp pass walks the chunks in order with the EXACT
// accumulator: when its binade matches a hypothesis and the chunk does not leave the binade, the chunk is
// applied in O(1); otherwise (the ~20 binade crossings, or a wrong guess) the chunk is staged in shared
// memory and added element by element with real FADDs.
enum { CHN_N = 512 };
struct ChainChunk { ParStep h0, h1; int e0; int n_unmarked; };
__device__ __forceinline__ void chain_bounds(int nb, int chunk, int* c0, int* c1) {
  const int per = (((nb + CHN_N - 1) / CHN_N) + 31) & ~31;
  *c0 = imin(chunk * per, nb); *c1 = imin(*c0 + per, nb);
}
__global__ void __launch_bounds__(256) k_chain_sums(const float* __restrict__ t, const int* __restrict__ status, double* __restrict__ csum, DevState* st) {
  const int nb = st->nb, lane = (int)threadIdx.x & 31;
  const int chunk = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (chunk >= CHN_N) return;
  int c0, c1; chain_bounds(nb, chunk, &c0, &c1);
  double s = 0;
  for (int i = c0 + lane; i < c1; i += 32) if (status[i] == 0) s += (double)t[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) csum[chunk] = s;
}
__global__ void __launch_bounds__(256) k_chain_compose(const float* __restrict__ t, const int* __restrict__ status, const double* __restrict__ csum,
                                                        ChainChunk* __restrict__ cc, DevState* st) {
  const int nb = st->nb, lane = (int)threadIdx.x & 31;
  const int chunk = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (chunk >= CHN_N) return;
  int c0, c1; chain_bounds(nb, chunk, &c0, &c1);
  double T0 = 0;
  for (int k = lane; k < chunk; k += 32) T0 += csum[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) T0 += __shfl_xor_sync(0xffffffffu, T0, o);
  // binade guess from the approximate total (the float chain drifts below / above the true sum by a few percent)
  const float g = (float)(T0 * 0.9);
  int e0 = g > 0.f ? (int)((__float_as_uint(g) >> 23) & 0xff) - 127 : -126;
  if (e0 < -100) e0 = -100;
  // lane-contiguous slices keep the order: lane l composes elements [a, b)
  const int per = (c1 - c0 + 31) / 32;
  const int a = imin(c0 + lane * per, c1), b = imin(a + per, c1);
  ParStep h[2]; int nun = 0;
#pragma unroll
  for (int hyp = 0; hyp < 2; ++hyp) {
    const int e = e0 + hyp;
    const double u = ldexp(1.0, e - 23), inv_u = ldexp(1.0, 23 - e);
    ParStep acc; acc.d0 = 0; acc.d1 = 0;
    for (int i = a; i < b; ++i) if (status[i] == 0) { acc = par_compose(acc, par_elem((double)t[i], u, inv_u)); if (hyp == 0) ++nun; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      ParStep up; up.d0 = __shfl_up_sync(0xffffffffu, acc.d0, o); up.d1 = __shfl_up_sync(0xffffffffu, acc.d1, o);
      if (lane >= o) acc = par_compose(up, acc);
    }
    h[hyp] = acc;   // lane 31 holds the whole chunk
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nun += __shfl_xor_sync(0xffffffffu, nun, o);
  if (lane == 31) { ChainChunk r; r.h0 = h[0]; r.h1 = h[1]; r.e0 = e0; r.n_unmarked = nun; cc[chunk] = r; }
}
__global__ void __launch_bounds__(32) k_chain_resolve(const float* __restrict__ t, const int* __restrict__ status, const ChainChunk* __restrict__ cc, DevState* st) {
  __shared__ float sv[4096];
  __shared__ ChainChunk scc[32];
  const int nb = st->nb, lane = (int)threadIdx.x;
  float S = 0.f; u32 n = 0;
  for (int cb = 0; cb < CHN_N; cb += 32) {
    scc[lane] = cc[cb + lane];
    __syncwarp();
    for (int k = 0; k < 32; ++k) {
      const int chunk = cb + k;
      int c0, c1; chain_bounds(nb, chunk, &c0, &c1);
      if (c0 >= c1) continue;                       // (uniform)
      int need = 1;
      if (lane == 0) {
        const ChainChunk r = scc[k];
        n += (u32)r.n_unmarked;
        const u32 sb = __float_as_uint(S);
        const int be = (int)((sb >> 23) & 0xff);      // biased exponent; normal positive floats only
        if (S > 0.f && be > 0) {
          const int e = be - 127;
          if (e == r.e0 || e == r.e0 + 1) {
            const ParStep hsel = e == r.e0 ? r.h0 : r.h1;
            const i64 N0 = (i64)((sb & 0x7fffffu) | 0x800000u);       // S = N0 * 2^(e-23), N0 in [2^23, 2^24)
            const i64 d = (N0 & 1) ? hsel.d1 : hsel.d0;
            const i64 N1 = N0 + d;
            if (N1 < 16777216) { S = __uint_as_float(((u32)be << 23) | ((u32)N1 & 0x7fffffu)); need = 0; }
          }
        }
      }
      need = __shfl_sync(0xffffffffu, need, 0);
      if (!need) continue;
      // in-order float adds for this chunk, staged through shared memory (marked bins contribute nothing)
      for (int s0 = c0; s0 < c1; s0 += 4096) {
        const int cnt = imin(4096, c1 - s0);
        for (int j = lane; j < cnt; j += 32) sv[j] = status[s0 + j] == 0 ? t[s0 + j] : -1.f;   // bin values are >= 0: -1 marks "skip"
        __syncwarp();
        if (lane == 0) {
          int j = 0;
          for (; j + 8 <= cnt; j += 8) {              // loads first, then the dependent FADD chain
            float x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = sv[j + q];
#pragma unroll
            for (int q = 0; q < 8; ++q) if (x[q] >= 0.f) S = __fadd_rn(S, x[q]);
          }
          for (; j < cnt; ++j) { const float x = sv[j]; if (x >= 0.f) S = __fadd_rn(S, x); }
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }
  if (lane == 0) { const int l0 = -st->st_lo; st->lvl_sum[l0] = S; st->lvl_cnt[l0] = n; }
}

// sums of the non-zero levels: one thread per level; the ordered (level, value) list of the marked
// bins is staged through shared memory in chunks that every thread scans (broadcast reads)
__global__ void __launch_bounds__(128) k_level_sums(const int* __restrict__ nz_lvl, const float* __restrict__ nz_val, DevState* st) {
  __shared__ int s_l[1024];
  __shared__ float s_v[1024];
  const int lo = st->st_lo, hi = st->st_hi, nz = st->n_nonzero;
  const int lvl = lo + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if ((int)(blockIdx.x * blockDim.x) + lo > hi) return;   // whole block beyond the last level
  float s = 0.f; u32 n = 0;
  for (int k0 = 0; k0 < nz; k0 += 1024) {
    const int cnt = imin(1024, nz - k0);
    __syncthreads();
    for (int k = (int)threadIdx.x; k < cnt; k += (int)blockDim.x) { s_l[k] = nz_lvl[k0 + k]; s_v[k] = nz_val[k0 + k]; }
    __syncthreads();
    for (int k = 0; k < cnt; ++k) if (s_l[k] == lvl) { s = __fadd_rn(s, s_v[k]); ++n; }
  }
  if (lvl <= hi && lvl != 0) { st->lvl_sum[lvl - lo] = s; st->lvl_cnt[lvl - lo] = n; }
}

__global__ void k_filter_params(DevState* st, int which) {
  if (threadIdx.x || blockIdx.x) return;
  const int lo = st->st_lo, hi = st->st_hi, nl = hi - lo + 1;
  const double dev = st->dev;
  st->filt_on = 0;
  for (int l = 0; l < nl; ++l) if (st->lvl_cnt[l] != 0) st->lvl_sum[l] = (float)((double)st->lvl_sum[l] / (double)st->lvl_cnt[l]);
  const float m0 = st->lvl_sum[-lo];
  int ldel = lo, ladd = hi;
  for (int l = 0; l < nl; ++l) if ((double)st->lvl_sum[l] < (double)m0 - dev) { ldel = l + lo; break; }
  for (int l = nl - 1; l >= 0; --l) if ((double)st->lvl_sum[l] > (double)m0 + dev) { ladd = l + lo; break; }
  {   // the level table as filterstatus prints it (rsi.cpp:991-997)
    RsiLogT& g = st->rlog[which];
    g.st_lo = lo; g.st_hi = hi; g.leveldel = ldel; g.leveladd = ladd;
    for (int l = 0; l < nl; ++l) { g.lvl_mean[l] = st->lvl_sum[l]; g.lvl_cnt[l] = st->lvl_cnt[l]; }
    g.filt_on = (ldel > 0 || ladd < 0 || ldel > ladd) ? 0 : 1;
  }
  if (ldel > 0 || ladd < 0 || ldel > ladd) return;
  st->filt_tdel = (double)m0 - dev; st->filt_tadd = (double)m0 + dev;
  st->filt_on = 1;
}

// the per-length counts and break levels of one rsistatus pass, kept for <out>.log (rsi.cpp:1221-1224, 1251-1254)
__global__ void k_rsi_log_pass(DevState* st, int which, int pass) {
  RsiLogT& g = st->rlog[which];
  for (int L = (int)threadIdx.x; L < LMAX_CAP + 2; L += (int)blockDim.x) { g.cnt[pass][0][L] = st->cnt_del[L]; g.cnt[pass][1][L] = st->cnt_dup[L]; }
  if (threadIdx.x == 0) { g.lbreak_del[pass] = st->lbreak_del; g.lbreak_dup[pass] = st->lbreak_dup; }
}

// trim both edges of every run but the last (rsi.cpp:1027-1046); one thread per run start
__global__ void k_filter_trim(const float* __restrict__ t, const int* __restrict__ sin, int* __restrict__ sout, DevState* st) {
  const int nb = st->nb, last = st->last_run_start;
  if (!st->filt_on) return;
  const double tdel = st->filt_tdel, tadd = st->filt_tadd;
  for (int j = (int)(blockIdx.x * blockDim.x + threadIdx.x); j < nb; j += (int)(gridDim.x * blockDim.x)) {
    const int s = sin[j];
    if (s == 0 || j == last) continue;
    if (j > 0 && same_run(sin[j - 1], s)) continue;
    int i1 = j, i2 = j;
    while (i2 + 1 < nb && same_run(sin[i2], sin[i2 + 1])) ++i2;
    while (((double)t[i1] > tdel && sin[i1] < 0) || ((double)t[i1] < tadd && sin[i1] > 0)) { sout[i1] = 0; ++i1; if (i1 >= i2) break; }
    // (the reference works in place; positions cleared by the first loop are never re-examined here
    //  except for a single-bin run, where clearing it twice changes nothing)
    while (((double)t[i2] > tdel && sin[i2] < 0) || ((double)t[i2] < tadd && sin[i2] > 0)) { sout[i2] = 0; --i2; if (i2 <= i1) break; }
  }
}

}  // namespace rsigpu
