// candidates.cuh -- the candidate stage of the rsi path as one cooperative thread block per contig.
//
// Replaces (reference file:line, relative to src/):
//   isitcnv            rsi.cpp:101-172     cnv_test_stats()
//   isitcnvwrap        rsi.cpp:175-287     cnv_test()            (neighbour walk: SURVEY.md A.1)
//   multisegments      rsi.cpp:368-410     SubsegIter
//   areblockscnv       rsi.cpp:415-546     blocks_test()
//   sortcnvstartposition rsi.cpp:549-577   list_sort()
//   optimize_with_derivative rsi.cpp:889-944  edge_refine()
//   mergesegments      rsi.cpp:694-885     merge_segments()
//   detectcnv tail     rsi.cpp:1860-1931   candidates_main()
//   expand_coordinate  rsi.cpp:1524-1551   expand_coord()
//   sd_filters         rsi.cpp:1753-1792   sd_filter_list()
//   partition_stat_tp  wufunctions.cpp:363-424  cta_hist_stat()
//   alglib::pnorm      alglib/specialfunctions.cpp:3152-3302  phi()
//
// Design: the reference walks neighbours one base at a time and copies whole call lists for every
// trial merge.  Here the walk is decomposed into "free" stretches (accepted = not an outlier) and
// "call zones" (first non-outlier triggers the jump), each handled with block-wide scans, so a
// candidate costs O(len / nthreads) steps; trial lists are overlays (two substituted entries)
// instead of copies; histogram quantiles are built with atomics and read with a chunked scan.
// Thread 0 owns all list mutation; every helper returns block-uniform values.
#pragma once
#include "../../include/rsigpu.h"
#include "cta.cuh"

namespace rsigpu {

typedef rsigpu_cnv Cnv;

enum { CAND_ERR_HIST = 1, CAND_ERR_REFCAP = 2, CAND_ERR_LISTCAP = 4, CAND_ERR_DEGENERATE = 8 };

struct CandCfg {
  int m, maxchkbp, merge, tid;
  double chklen, minmlen, buffer, p;
  double rdmedian;  // rsi::RDmedian at the time of the call
  double rdsd;      // rsi::RDsd
  int span;         // rsi::end - rsi::start + 1 (length of the per-base array after N removal)
};

struct CandScratch {
  int* ref;        int ref_cap;   // gathered neighbours
  int* sub;        int sub_cap;   // sub-sampled ref + cnv (>= maxchkbp*10 + 2)
  long long* pref;                // ref_cap + 1 prefix sums
  float* rm;                      // ref_cap running means
  unsigned* hist;  int hist_cap;  // histogram buckets (global memory fallback)
  unsigned* shist; int shist_cap; // histogram buckets in shared memory (preferred when they fit)
  int* err;                       // sticky error bits (CAND_ERR_*)
  long long* prof;                // optional per-phase clock64 totals (thread 0), 16 slots, may be null
};

#if defined(__CUDACC__)
RSI_DEV long long cand_clock() { return clock64(); }
#else
inline long long cand_clock() { return 0; }
#endif
// phase timer: thread 0 adds the elapsed clocks since *t0 to slot k and restarts the timer
RSI_DEV void cand_tick(const Cta& c, const CandScratch& S, int k, long long* t0) {
  if (S.prof && c.tid == 0) { const long long t = cand_clock(); S.prof[k] += t - *t0; *t0 = t; }
}

RSI_DEV Cnv cnv_default() {  // cnv_st(), rsi.h:30-50
  Cnv c;
  c.tid = -1; c.type = RSIGPU_TYPE_UNKNOWN; c.geno = 0; c.status = 0; c.start = 0; c.end = 0; c.length = 0;
  c.sc1 = 0; c.sc2 = 0; c.pair = 0; c.score = 0.0; c.p1 = 1.0; c.p2 = 1.0; c.cnvmed = 0; c.cnvsd = 0; c.cnviqr = 0;
  c.refmed = 0; c.refsd = 0; c.refiqr = 0; c.q0 = -1.0; c.rp = -1; c.pad_ = 0;
  return c;
}

// ---------------------------------------------------------------------------------------------
// Phi(x) = (1 + erf(x / sqrt 2)) / 2 with the Cephes rational approximations ALGLIB uses.
RSI_DEV double horner(const double* k, int n, double x) {  // k[0] is the leading coefficient
  double r = 0.0;
  for (int i = 0; i < n; ++i) r = k[i] + x * r;
  return r;
}
RSI_DEV double erfc_tail(double x) {  // 0.5 <= x < 10
  const double P[8] = {0.5641877825507397413087057563, 9.675807882987265400604202961, 77.08161730368428609781633646,
                       368.5196154710010637133875746, 1143.262070703886173606073338, 2320.439590251635247384768711,
                       2898.0293292167655611275846, 1826.3348842295112592168999};
  const double Q[9] = {1.0, 17.14980943627607849376131193, 137.1255960500622202878443578, 661.7361207107653469211984771,
                       2094.384367789539593790281779, 4429.612803883682726711528526, 6089.5424232724435504633068,
                       4958.82756472114071495438422, 1826.3348842295112595576438};
  return exp(-(x * x)) * horner(P, 8, x) / horner(Q, 9, x);
}
RSI_DEV double erf_small(double x) {  // 0 <= x < 0.5
  const double P[7] = {0.007547728033418631287834, -0.288805137207594084924010, 14.3383842191748205576712,
                       38.0140318123903008244444, 3017.82788536507577809226, 7404.07142710151470082064,
                       80437.3630960840172832162};
  const double Q[6] = {1.00000000000000000000000, 38.0190713951939403753468, 658.070155459240506326937,
                       6379.60017324428279487120, 34216.5257924628539769006, 80437.3630960840172826266};
  double xsq = x * x;
  return 1.1283791670955125738961589031 * x * horner(P, 7, xsq) / horner(Q, 6, xsq);
}
RSI_DEV double phi(double v) {
  double z = v / 1.41421356237309504880;
  double s = z > 0 ? 1.0 : (z < 0 ? -1.0 : 0.0);
  double a = fabs(z), e;
  if (a < 0.5) e = s * erf_small(a);
  else if (a >= 10) e = s;
  else e = s * (1 - erfc_tail(a));
  return 0.5 * (e + 1);
}

// ---------------------------------------------------------------------------------------------
// Histogram quartiles/median of x[0..n): bucket (x-ymin)/dy+0.5, value ymin+b*dy at the first bucket
// whose running count reaches n/4, n/2, 3n/4; mean/min/max when the range is below dy.
template <class T>
RSI_DEVN void cta_hist_stat(const Cta& c, const CandScratch& S, const T* x, int n, double dy, double q[3]) {
  double mn = 1e300, mx = -1e300, sm = 0.0;
  {
    int i = c.tid;
    for (; i + 3 * c.nthr < n; i += 4 * c.nthr) {   // four independent loads in flight per thread
      const T a0 = x[i], a1 = x[i + c.nthr], a2 = x[i + 2 * c.nthr], a3 = x[i + 3 * c.nthr];
      const double v0 = (double)a0, v1 = (double)a1, v2 = (double)a2, v3 = (double)a3;
      mn = v0 < mn ? v0 : mn; mx = v0 > mx ? v0 : mx; mn = v1 < mn ? v1 : mn; mx = v1 > mx ? v1 : mx;
      mn = v2 < mn ? v2 : mn; mx = v2 > mx ? v2 : mx; mn = v3 < mn ? v3 : mn; mx = v3 > mx ? v3 : mx;
      sm += v0; sm += v1; sm += v2; sm += v3;
    }
    for (; i < n; i += c.nthr) { double v = (double)x[i]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; sm += v; }
  }
  mn = c.reduce_ol(mn, MinOp()); mx = c.reduce_ol(mx, MaxOp()); sm = c.reduce_ol(sm, SumOp());
  q[0] = mn; q[1] = sm / (double)n; q[2] = mx;
  if ((mx - mn) < dy) return;
  const size_t np = (size_t)((mx - mn) / dy + 2);
  unsigned* H = (S.shist && np + 1 <= (size_t)S.shist_cap) ? S.shist : S.hist;
  if (H == S.hist && np + 1 > (size_t)S.hist_cap) { if (c.tid == 0) *S.err |= CAND_ERR_HIST; return; }
  for (size_t b = c.tid; b <= np; b += c.nthr) H[b] = 0u;
  c.sync();
  // neighbouring samples mostly share a bucket: one atomic per distinct bucket of a warp
  for (int base = 0; base < n; base += 4 * c.nthr) {
    T a[4]; bool valid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int i = base + k * c.nthr + c.tid; valid[k] = i < n; a[k] = valid[k] ? x[i] : T(0); }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      size_t b = 0;
      if (valid[k]) {
        if (dy == 1.0) b = (size_t)((double)a[k] - mn + 0.5);       // integer samples: the division by dy = 1 is exact anyway
        else {
          // bucket = (size_t)((x - mn) / dy + 0.5).  dy = 0.01: multiply by 100 (error << 1e-6 for < 2^22 buckets) and fall back
          // to the exact IEEE division only when the result sits within 1e-6 of a bucket edge.
          const double t = (double)a[k] - mn;
          double qv = t * 100.0 + 0.5, fl = floor(qv);
          if (dy != 0.01 || qv - fl < 1e-6 || fl + 1.0 - qv < 1e-6) { qv = t / dy + 0.5; fl = floor(qv); }
          b = (size_t)fl;
        }
      }
      cta_hist_add(H, b, valid[k]);
    }
  }
  c.sync();
  const unsigned long long r4 = (unsigned long long)n / 4, r2 = (unsigned long long)n / 2, r34 = (unsigned long long)n * 3 / 4;
  const size_t chunk = (np + c.nthr - 1) / c.nthr;
  const size_t b0 = (size_t)c.tid * chunk < np ? (size_t)c.tid * chunk : np, b1 = (b0 + chunk < np) ? b0 + chunk : np;
  long long local = 0;
  for (size_t b = b0; b < b1; ++b) local += H[b];
  long long tot;
  long long run = c.scan_excl_ol(local, &tot);
  if (c.tid == 0) { c.bc[8] = q[0]; c.bc[9] = q[1]; c.bc[10] = q[2]; }
  c.sync();
  for (size_t b = b0; b < b1; ++b) {
    unsigned long long cnt = H[b], before = (unsigned long long)run, after = before + cnt;
    if (before < r4 && after >= r4) c.bc[8] = mn + b * dy;
    if (before < r2 && after >= r2) c.bc[9] = mn + b * dy;
    if (before < r34 && after >= r34) c.bc[10] = mn + b * dy;
    run += (long long)cnt;
  }
  c.sync();
  q[0] = c.bc[8]; q[1] = c.bc[9]; q[2] = c.bc[10];
  c.sync();
}

// ---------------------------------------------------------------------------------------------
// A call list with up to two substituted entries (what the reference builds by copying the list).
struct ListView {
  const Cnv* base; int n;
  int oi0; const Cnv* o0;
  int oi1; const Cnv* o1;
  RSI_DEV const Cnv& at(int j) const { return j == oi0 ? *o0 : (j == oi1 ? *o1 : base[j]); }
};
RSI_DEV ListView plain_view(const Cnv* base, int n) { ListView v; v.base = base; v.n = n; v.oi0 = -1; v.o0 = nullptr; v.oi1 = -1; v.o1 = nullptr; return v; }

RSI_DEV bool nb_accept(int v, int flag, double up, double lo) {
  if (flag == RSIGPU_TYPE_DEL && v > up) return false;
  if (flag == RSIGPU_TYPE_DUP && v < lo) return false;
  return true;
}

// Scan positions lo..hi (dir=+1) or hi..lo (dir=-1); store the first `want` accepted values to
// dst[0..), in scan order.  Returns how many were stored (block-uniform).
#if defined(RSI_CTA_PARALLEL)
// Every WARP takes a contiguous segment of a super-chunk sized to what is still wanted; lanes read
// stride-1 (coalesced), __ballot_sync + popc give the compaction offsets, one exchange of the 32 warp
// counts per super-chunk.
// Sum of the per-warp slots (16-byte stride, one value per warp of the whole group, in a cluster: global memory) before
// `warp` and over all `nw` warps: the lanes of a warp read 32 slots at a time and add up by shuffles, instead of every
// thread reading every slot (up to 256 dependent-latency loads per thread in an 8-CTA cluster).
template <class T>
RSI_DEV void cta_slot_sums(const unsigned char* slots, int warp, int nw, int lane, T* before, T* total) {
  T b = 0, t = 0;
  for (int w0 = 0; w0 < nw; w0 += 32) {
    const int w = w0 + lane;
    const T v = w < nw ? *reinterpret_cast<const volatile T*>(slots + 16 * w) : (T)0;
    t += v; if (w < warp) b += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { b += __shfl_xor_sync(0xffffffffu, b, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
  *before = b; *total = t;
}
RSI_DEVN int cta_collect(const Cta& c, const int* RD, int lo, int hi, int dir, int flag, double up, double lw, int want, int* dst) {
  int got = 0, pos = 0;
  const int len = hi - lo + 1;
  const int lane = c.tid & 31, warp = c.tid >> 5, nw = c.nthr >> 5;
  unsigned char* slots = c.wslots();
  while (pos < len && got < want) {
    int sc = want - got + c.nthr;               // nearly every position is accepted
    if (sc > len - pos) sc = len - pos;
    const int seg = (((sc + nw - 1) / nw) + 31) & ~31;
    const int w0 = pos + (warp * seg < sc ? warp * seg : sc), w1 = pos + ((warp + 1) * seg < sc ? (warp + 1) * seg : sc);
    int cnt = 0;
    for (int j0 = w0; j0 < w1; j0 += 32) {
      const int j = j0 + lane;
      const bool ok = j < w1 && nb_accept(RD[dir > 0 ? lo + j : hi - j], flag, up, lw);
      cnt += __popc(__ballot_sync(0xffffffffu, ok));
    }
    c.sync();
    if (lane == 0) *reinterpret_cast<int*>(slots + 16 * warp) = cnt;
    c.sync();
    int base = 0, tot = 0;
    cta_slot_sums<int>(slots, warp, nw, lane, &base, &tot);
    int off = got + base;
    for (int j0 = w0; j0 < w1 && off < want; j0 += 32) {
      const int j = j0 + lane;
      int v = 0; bool ok = false;
      if (j < w1) { v = RD[dir > 0 ? lo + j : hi - j]; ok = nb_accept(v, flag, up, lw); }
      const unsigned bm = __ballot_sync(0xffffffffu, ok);
      const int my = off + __popc(bm & ((1u << lane) - 1u));
      if (ok && my < want) dst[my] = v;
      off += __popc(bm);
    }
    got = got + tot < want ? got + tot : want;
    pos += sc;
  }
  c.sync();
  return got;
}
#else
RSI_DEVN int cta_collect(const Cta& c, const int* RD, int lo, int hi, int dir, int flag, double up, double lw, int want, int* dst) {
  int got = 0;
  const int len = hi - lo + 1;
  for (int j = 0; j < len && got < want; ++j) { const int v = RD[dir > 0 ? lo + j : hi - j]; if (nb_accept(v, flag, up, lw)) dst[got++] = v; }
  (void)c;
  return got;
}
#endif
// First accepted position in scan order, or -1 (block-uniform).
RSI_DEVN int cta_find_first(const Cta& c, const int* RD, int lo, int hi, int dir, int flag, double up, double lw) {
  const int len = hi - lo + 1;
  for (int base = 0; base < len; base += c.nthr * 8) {
    int best = 0x7fffffff;
    for (int k = 0; k < 8; ++k) {
      const int j = base + k * c.nthr + c.tid;
      if (j < len && j < best && nb_accept(RD[dir > 0 ? lo + j : hi - j], flag, up, lw)) best = j;
    }
    best = c.reduce_ol(best, MinOp());
    if (best != 0x7fffffff) return dir > 0 ? lo + best : hi - best;
  }
  return -1;
}

// pref[0] = 0, pref[j+1] = ref[0] + .. + ref[j]  (exact integer prefix sums)
#if defined(RSI_CTA_PARALLEL)
RSI_DEVN void cta_prefix_i32(const Cta& c, const int* ref, int n, long long* pref) {
  const int lane = c.tid & 31, warp = c.tid >> 5, nw = c.nthr >> 5;
  const int seg = (((n + nw - 1) / nw) + 31) & ~31;
  const int w0 = warp * seg < n ? warp * seg : n, w1 = (warp + 1) * seg < n ? (warp + 1) * seg : n;
  long long loc = 0;
#pragma unroll 4
  for (int j = w0 + lane; j < w1; j += 32) loc += (long long)ref[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
  unsigned char* slots = c.wslots();
  c.sync();
  if (lane == 0) *reinterpret_cast<long long*>(slots + 16 * warp) = loc;
  c.sync();
  long long carry = 0, all_ = 0;
  cta_slot_sums<long long>(slots, warp, nw, lane, &carry, &all_);
  if (c.tid == 0) pref[0] = 0;
  for (int j0 = w0; j0 < w1; j0 += 32) {
    const int j = j0 + lane;
    long long inc = j < w1 ? (long long)ref[j] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (j < w1) pref[j + 1] = carry + inc;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  c.sync();
}
#else
RSI_DEVN void cta_prefix_i32(const Cta& c, const int* ref, int n, long long* pref) {
  long long run = 0; pref[0] = 0;
  for (int j = 0; j < n; ++j) { run += ref[j]; pref[j + 1] = run; }
  (void)c;
}
#endif

// isitcnv: statistics of the candidate against the running mean of its neighbours; fills *out.
RSI_DEVN void cnv_test_stats(const Cta& c, const CandCfg& P, const CandScratch& S, const int* ref, int nref, const int* cnv, int ncnv, Cnv* out) {
  const int d = ncnv, nr = nref - d;
  long long t0 = cand_clock();
  if (nr <= 0 || d <= 0) {  // the reference aborts here (Array bounds throw)
    if (c.tid == 0) { out->status = -9; out->geno = 0; *S.err |= CAND_ERR_DEGENERATE; }
    c.sync();
    return;
  }
  // exact integer prefix sums of the neighbours (the reference slides a double sum of ints: exact)
  cta_prefix_i32(c, ref, nref, S.pref);
  c.sync();
  cand_tick(c, S, 4, &t0);
  double s1 = 0.0, s2 = 0.0;
  {
    const long long* __restrict__ pf = S.pref;
    float* __restrict__ rmo = S.rm;
    int i = c.tid;
    for (; i + 3 * c.nthr < nr; i += 4 * c.nthr) {
      long long w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) w[k] = pf[i + k * c.nthr + d] - pf[i + k * c.nthr];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float mnv = (float)((double)w[k] / (double)d);
        rmo[i + k * c.nthr] = mnv;
        s1 += (double)mnv; s2 += (double)mnv * (double)mnv;
      }
    }
    for (; i < nr; i += c.nthr) {
      float mnv = (float)((double)(pf[i + d] - pf[i]) / (double)d);
      rmo[i] = mnv;
      s1 += (double)mnv; s2 += (double)mnv * (double)mnv;
    }
  }
  c.sync();
  s1 = c.reduce_ol(s1, SumOp()); s2 = c.reduce_ol(s2, SumOp());
  cand_tick(c, S, 5, &t0);
  double rq[3], cq[3];
  cta_hist_stat(c, S, S.rm, nr, 0.01, rq);
  cand_tick(c, S, 6, &t0);
  cta_hist_stat(c, S, cnv, ncnv, 1.0, cq);
  cand_tick(c, S, 7, &t0);
  long long a1 = 0, a2 = 0;
#pragma unroll 4
  for (int i = c.tid; i < ncnv; i += c.nthr) { long long v = cnv[i]; a1 += v; a2 += v * v; }
  a1 = c.reduce_ol(a1, SumOp()); a2 = c.reduce_ol(a2, SumOp());
  cand_tick(c, S, 8, &t0);
  if (c.tid == 0) {
    double rmean = s1 / double(nr);
    double rsd = sqrt(s2 / double(nr) - rmean * rmean);
    const double rmed = rq[1];
    if (rsd < 1E-3) rsd = rmed / 40.0 + 1E-3;
    double cmean = (double)a1 / double(ncnv);
    out->length = out->end - out->start + 1;
    out->cnvmed = cq[1];
    out->cnvsd = sqrt((double)a2 / double(ncnv) - cmean * cmean);
    out->cnviqr = cq[2] - cq[0];
    out->refmed = rmed;
    out->refiqr = rq[2] - rq[0];
    out->refsd = out->refiqr / 1.349;
    out->geno = 1; out->status = 1;
    int flag = out->cnvmed > P.rdmedian ? RSIGPU_TYPE_DUP : RSIGPU_TYPE_DEL;
    if (out->type == RSIGPU_TYPE_UNKNOWN) out->type = flag;
    if (out->type != flag) out->status = -9;  // "basic assignment error": geno stays 1
    else if (out->type == RSIGPU_TYPE_DEL) {
      double reference = rmed < P.rdmedian ? rmed : P.rdmedian;
      if (reference < 0.8 * P.rdmedian) reference = 0.8 * P.rdmedian;
      double nu = (3.0 * out->cnvmed - 2.0 * reference) / rsd;
      out->p1 = phi(nu);
      if (nu > 0) { out->status = -9; out->geno = 0; }
    } else {
      double reference = rmed > P.rdmedian ? rmed : P.rdmedian;
      double nu = (2.5 * out->cnvmed - 3.0 * reference) / rsd / 1.5;
      out->p1 = 1.0 - phi(nu);
      if (nu < 0) { out->status = -9; out->geno = 0; }
    }
  }
  c.sync();
}

// isitcnvwrap: gather <= chklen*d accepted neighbours per side around entry `ci` of the (overlaid)
// list, then run the test.  `out` is the entry being tested (view.at(ci) must alias it).
RSI_DEVN void cnv_test(const Cta& c, const CandCfg& P, const CandScratch& S, const int* RD, int n, const ListView& V, int ci, Cnv* out) {
  const Cnv& me = V.at(ci);
  const int flag = me.type, start = me.start, end = me.end;
  const int cnvlen = end - start + 1, nl = V.n;
  const int pts = P.maxchkbp * 10;
  int d = cnvlen;
  if (n == P.span) { if (d < P.m * P.minmlen) d = (int)(P.m * P.minmlen); }
  if (n < P.span / 2) { if (d < P.minmlen) d = (int)P.minmlen + 1; }
  const int refsize = (int)(P.chklen * d * 2);
  const int buffer = (int)(cnvlen * P.buffer + 1);
  const double up = P.rdmedian * 3.0, lw = P.rdmedian * 0.15;

  // ---- left side, nearest first, then reversed into ascending position order
  long long t0 = cand_clock();
  if (S.prof && c.tid == 0) { S.prof[0] += 1; S.prof[14] += cnvlen; }
  int i = start - buffer, idx = ci - 1;
  while (i > 0 && idx > 0 && i < V.at(idx).start) --idx;
  while (idx > 0 && V.at(idx).status == -9) --idx;
  int k = (int)(P.chklen * d - 1);
  if (n - end < P.chklen * d) k = refsize - 1 - n + end;
  int nleft = 0;
  const int left_want = k + 1;
  bool overflow = false;
  while (i > 2 && nleft < left_want) {
    int lo = 2, zone = 0;
    if (idx >= 0) {
      const int s = V.at(idx).start, e = V.at(idx).end;
      if (s > i - 1) { idx = -1; continue; }       // call lies right of the walk: it can never match again
      if (e < i - 1) lo = e + 1 > 2 ? e + 1 : 2;   // free stretch down to the call's end
      else { zone = 1; lo = s > 2 ? s : 2; }       // inside the call: first non-outlier triggers the jump
    }
    if (!zone) {
      int want = left_want - nleft;
      if (nleft + (want < i - lo ? want : i - lo) > S.ref_cap) { overflow = true; break; }
      nleft += cta_collect(c, RD, lo, i - 1, -1, flag, up, lw, want, S.ref + nleft);
      i = lo;
    } else {
      int f = cta_find_first(c, RD, lo, i - 1, -1, flag, up, lw);
      if (f >= 0) { i = V.at(idx).start - 1; --idx; while (idx > 0 && V.at(idx).status == -9) --idx; }
      else { i = lo; idx = -1; }
    }
  }
  cand_tick(c, S, 1, &t0);
  for (int j = c.tid; j < nleft / 2; j += c.nthr) { int a = S.ref[j], b = S.ref[nleft - 1 - j]; S.ref[j] = b; S.ref[nleft - 1 - j] = a; }
  c.sync();
  cand_tick(c, S, 2, &t0);

  // ---- right side, ascending
  int kk = nleft;
  i = end + buffer; idx = ci + 1;
  while (i < n - 2 && idx < nl && i > V.at(idx).end) ++idx;
  while (idx < nl - 1 && V.at(idx).status == -9) ++idx;
  while (!overflow && i < n - 2 && kk < refsize) {
    int hi = n - 2, zone = 0;
    if (idx < nl) {
      const int s = V.at(idx).start, e = V.at(idx).end;
      if (e < i + 1) { idx = nl; continue; }       // call lies left of the walk
      if (s > i + 1) hi = s - 1 < n - 2 ? s - 1 : n - 2;
      else { zone = 1; hi = e < n - 2 ? e : n - 2; }
    }
    if (!zone) {
      int want = refsize - kk;
      if (kk + (want < hi - i ? want : hi - i) > S.ref_cap) { overflow = true; break; }
      kk += cta_collect(c, RD, i + 1, hi, +1, flag, up, lw, want, S.ref + kk);
      i = hi;
    } else {
      int f = cta_find_first(c, RD, i + 1, hi, +1, flag, up, lw);
      if (f >= 0) { i = V.at(idx).end + 1; ++idx; while (idx < nl - 1 && V.at(idx).status == -9) ++idx; }
      else { i = hi; idx = nl; }
    }
  }
  if (overflow) {
    if (c.tid == 0) { *S.err |= CAND_ERR_REFCAP; out->status = -9; out->geno = 0; }
    c.sync();
    return;
  }
  cand_tick(c, S, 3, &t0);
  if (S.prof && c.tid == 0) S.prof[15] += kk;
  const int* ref = S.ref; int nref = kk;
  const int* cnv = RD + start; int ncnv = cnvlen;
  if (nref + ncnv > pts) {  // sub-sample both to ~pts points, rsi.cpp:264-282
    const int tot = nref + ncnv;
    const int dref = (int)((double)nref / (double)tot * (double)pts);
    const int dcnv = (int)((double)ncnv / (double)tot * (double)pts);
    if (dref + dcnv > S.sub_cap) { if (c.tid == 0) { *S.err |= CAND_ERR_REFCAP; out->status = -9; out->geno = 0; } c.sync(); return; }
    for (int a = c.tid; a < dcnv; a += c.nthr) S.sub[dref + a] = cnv[(int)(double(a) / double(dcnv) * double(ncnv))];
    for (int a = c.tid; a < dref; a += c.nthr) S.sub[a] = ref[(int)(double(a) / double(dref) * double(nref))];
    c.sync();
    for (int a = c.tid; a < dref; a += c.nthr) S.ref[a] = S.sub[a];
    c.sync();
    nref = dref; cnv = S.sub + dref; ncnv = dcnv;
  }
  cnv_test_stats(c, P, S, ref, nref, cnv, ncnv, out);
}

// ---------------------------------------------------------------------------------------------
// Stable sort by start (after un-reversing entries): rank sort, O(n^2 / nthreads).
RSI_DEVN void list_sort(const Cta& c, Cnv* list, int n, Cnv* tmp) {
  for (int j = c.tid; j < n; j += c.nthr) if (list[j].start > list[j].end) { int t = list[j].start; list[j].start = list[j].end; list[j].end = t; }
  c.sync();
  for (int j = c.tid; j < n; j += c.nthr) {
    int r = 0; const int sj = list[j].start;
    for (int t = 0; t < n; ++t) { int st = list[t].start; r += (st < sj || (st == sj && t < j)) ? 1 : 0; }
    tmp[r] = list[j];
  }
  c.sync();
  for (int j = c.tid; j < n; j += c.nthr) list[j] = tmp[j];
  c.sync();
}

// optimize_with_derivative for one call: dd(i) = sum RD[i-len,i) - sum RD[i,i+len) for i in
// [nstart,nend); start -> first arg-max (DEL) / arg-min (DUP) over the first 2*disp values, end ->
// the opposite over the last 2*disp; an extremum at offset 0 (or none beyond 0) changes nothing.
RSI_DEVN void edge_refine(const Cta& c, const int* RD, int n, Cnv* cv) {
  const int len = cv->end - cv->start + 1;
  const int disp = 250 > len / 4 ? 250 : len / 4;
  const int ns = cv->start - disp, ne = cv->end + disp;
  const int type = cv->type;
  c.sync();
  if (ns < 2 * len || ne > n - 2 * len) return;
  const int ndd = ne - ns;
  long long a = 0;
  for (int j = c.tid; j < len; j += c.nthr) a += (long long)RD[ns - len + j] - (long long)RD[ns + j];
  const long long dd0 = c.reduce_ol(a, SumOp());
  const int tail0 = ndd - 2 * disp;
  ValIdx headbest; headbest.v = 0.0; headbest.i = -1;   // running extremum over the head window
  ValIdx tailbest; tailbest.v = 0.0; tailbest.i = -1;
  const double hs = type == RSIGPU_TYPE_DEL ? 1.0 : -1.0;  // DEL: head max / tail min; DUP: head min / tail max
  {
    // dd(j) = dd(0) + sum_{1<=k<=j} inc(k); a warp owns a contiguous slice of j, lanes stride-1
    const bool typed = type == RSIGPU_TYPE_DEL || type == RSIGPU_TYPE_DUP;
#if defined(RSI_CTA_PARALLEL)
    const int lane = c.tid & 31, warp = c.tid >> 5, nw = c.nthr >> 5;
    const int seg = (((ndd + nw - 1) / nw) + 31) & ~31;
    const int w0 = warp * seg < ndd ? warp * seg : ndd, w1 = (warp + 1) * seg < ndd ? (warp + 1) * seg : ndd;
    long long loc = 0;
#pragma unroll 2
    for (int j = w0 + lane; j < w1; j += 32) if (j >= 1) { const int p = ns + j; loc += -(long long)RD[p - 1 - len] + 2ll * RD[p - 1] - (long long)RD[p - 1 + len]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
    unsigned char* slots = c.wslots();
    c.sync();
    if (lane == 0) *reinterpret_cast<long long*>(slots + 16 * warp) = loc;
    c.sync();
    long long carry = 0, all_ = 0;
    cta_slot_sums<long long>(slots, warp, nw, lane, &carry, &all_);
    carry += dd0;
    for (int j0 = w0; j0 < w1; j0 += 32) {
      const int j = j0 + lane;
      long long inc = 0;
      if (j < w1 && j >= 1) { const int p = ns + j; inc = -(long long)RD[p - 1 - len] + 2ll * RD[p - 1] - (long long)RD[p - 1 + len]; }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
      if (j < w1 && typed) {
        const double ddj = (double)(carry + inc);
        if (j < 2 * disp && hs * ddj > headbest.v) { headbest.v = hs * ddj; headbest.i = j; }
        if (j >= tail0 && -hs * ddj > tailbest.v) { tailbest.v = -hs * ddj; tailbest.i = j; }
      }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    c.sync();
#else
    long long run = dd0;
    for (int j = 0; j < ndd; ++j) {
      if (j >= 1) { const int p = ns + j; run += -(long long)RD[p - 1 - len] + 2ll * RD[p - 1] - (long long)RD[p - 1 + len]; }
      const double ddj = (double)run;
      if (typed) {
        if (j < 2 * disp && hs * ddj > headbest.v) { headbest.v = hs * ddj; headbest.i = j; }
        if (j >= tail0 && -hs * ddj > tailbest.v) { tailbest.v = -hs * ddj; tailbest.i = j; }
      }
    }
#endif
  }
  // per-thread candidates are each thread's first strict maximum in ascending j; merge with first-index ties
  if (headbest.i < 0) { headbest.v = 0.0; headbest.i = 0x7fffffffffffll; }
  if (tailbest.i < 0) { tailbest.v = 0.0; tailbest.i = 0x7fffffffffffll; }
  headbest = c.reduce_ol(headbest, ArgMaxFirst());
  tailbest = c.reduce_ol(tailbest, ArgMaxFirst());
  c.sync();
  if (c.tid == 0) {
    if (headbest.v > 0.0 && headbest.i > 0 && headbest.i < ndd) cv->start = ns + (int)headbest.i;
    if (tailbest.v > 0.0 && tailbest.i > 0 && tailbest.i < ndd) cv->end = ne - ndd + (int)tailbest.i;
  }
  c.sync();
}

RSI_DEVN double cta_mean(const Cta& c, const int* RD, int lo, int hi) {  // mean_tp: exact integer sum / count
  long long a = 0;
  for (int j = lo + c.tid; j <= hi; j += c.nthr) a += RD[j];
  a = c.reduce_ol(a, SumOp());
  return (double)a / double(hi - lo + 1);
}

// remove status == -9 entries in place, keeping order (thread 0)
RSI_DEV int list_drop_deleted(const Cta& c, Cnv* list, int n) {
  int w = n;
  c.sync();
  if (c.tid == 0) { w = 0; for (int j = 0; j < n; ++j) if (list[j].status != -9) { if (w != j) list[w] = list[j]; ++w; } }
  return cta_bcast(c, w, 0);
}

// mergesegments.  `ov` = two scratch entries in global memory for the overlays.
RSI_DEVN int merge_segments(const Cta& c, const CandCfg& P, const CandScratch& S, const int* RD, int n, Cnv* list, int nl, Cnv* ov) {
  Cnv* A = ov; Cnv* B = ov + 1;
  for (int i = 0; i < nl - 1; ++i) {
    c.sync();
    const Cnv x = list[i], y = list[i + 1];
    if (x.type != y.type) continue;
    const int mxs = x.start > y.start ? x.start : y.start, mne = x.end < y.end ? x.end : y.end;
    if (!(mxs < mne)) continue;
    if (c.tid == 0) {
      *A = x; A->start = x.start < y.start ? x.start : y.start; A->end = x.end > y.end ? x.end : y.end;
      *B = *A; B->status = -9;
    }
    c.sync();
    ListView V = plain_view(list, nl); V.oi0 = i; V.o0 = A; V.oi1 = i + 1; V.o1 = B;
    cnv_test(c, P, S, RD, n, V, i, A);
    int geno = A->geno;
    c.sync();
    if (geno == 0) {
      if (c.tid == 0) { *A = x; *B = y; B->status = -9; }
      c.sync();
      cnv_test(c, P, S, RD, n, V, i, A);
      if (c.tid == 0) { A->status = -9; B->status = 0; }
      c.sync();
      cnv_test(c, P, S, RD, n, V, i + 1, B);
      if (c.tid == 0) {
        if (B->p1 < A->p1) *A = *B;
        if (A->p1 > P.p) { list[i].status = -9; list[i + 1].status = -9; }
      }
      c.sync();
      geno = A->geno;
      c.sync();
    }
    if (geno == 0) continue;
    if (c.tid == 0) { list[i] = *A; list[i].status = -9; list[i + 1] = *A; list[i + 1].status = 0; }
    c.sync();
  }
  nl = list_drop_deleted(c, list, nl);
  if (!P.merge) return nl;
  for (int i = 0; i < nl - 1; ++i) {
    c.sync();
    const Cnv x = list[i], y = list[i + 1];
    if (x.type != y.type) continue;
    if (x.geno == 0 || y.geno == 0) continue;
    const int gap = y.start - x.end;
    if (gap > (x.end - x.start) * P.chklen * 0.7 && gap > (y.end - y.start) * P.chklen * 0.7) continue;
    const double m1 = cta_mean(c, RD, x.start, x.end), m2 = cta_mean(c, RD, y.start, y.end);
    const double cm = (m1 * (x.end - x.start) + m2 * (y.end - y.start)) / ((x.end - x.start) + (y.end - y.start));
    const double mm = cta_mean(c, RD, x.start, y.end);
    if (x.type == RSIGPU_TYPE_DEL && mm > cm + 1.5 * y.refsd + 1.5 * x.refsd) continue;
    if (x.type == RSIGPU_TYPE_DUP && mm < cm - 1.5 * y.refsd - 1.5 * x.refsd) continue;
    if (c.tid == 0) { *A = x; A->end = y.end; *B = *A; B->status = -9; }
    c.sync();
    ListView V = plain_view(list, nl); V.oi0 = i; V.o0 = A; V.oi1 = i + 1; V.o1 = B;
    cnv_test(c, P, S, RD, n, V, i, A);
    const int geno = A->geno;
    c.sync();
    if (geno == 0) continue;
    if (c.tid == 0) { list[i] = *A; list[i + 1] = *A; list[i].status = -9; }
    c.sync();
  }
  return list_drop_deleted(c, list, nl);
}

// multisegments as a generator: nested sub-level runs of a failed segment, in the reference's
// order (level ascending, position ascending), the last run of every level dropped.
struct SubsegIter {
  int seg_start, len, lo, hi, level, pos, open, s0, s1;
};
RSI_DEV void subseg_begin(SubsegIter& it, const Cnv& seg, const int* status) {
  it.seg_start = seg.start; it.len = seg.end - seg.start + 1;
  const int* s2 = status + seg.start;
  it.lo = s2[0]; it.hi = s2[0];
  for (int i = 0; i < it.len; ++i) { it.lo = s2[i] < it.lo ? s2[i] : it.lo; it.hi = s2[i] > it.hi ? s2[i] : it.hi; }
  it.level = it.lo; it.pos = 0; it.open = 0; it.s0 = it.s1 = 0;
}
RSI_DEV bool subseg_in_level(int v, int lev) { return v != 0 && ((lev < 0 && v < 0 && v >= lev) || (lev > 0 && v > 0 && v <= lev)); }
// thread 0 only; returns true and fills (a,b) (absolute bin coordinates) for the next sub-segment
RSI_DEV bool subseg_next(SubsegIter& it, const int* status, int* a, int* b) {
  const int* s2 = status + it.seg_start;
  while (it.level < it.hi) {
    if (it.level == 0) { ++it.level; it.pos = 0; it.open = 0; continue; }
    if (it.pos == 0 && !it.open) {  // levelcount test
      int hits = 0;
      for (int i = 0; i < it.len; ++i) hits += s2[i] == it.level;
      if (hits == 0) { ++it.level; continue; }
    }
    while (it.pos < it.len) {
      int i = it.pos++;
      if (!subseg_in_level(s2[i], it.level)) continue;
      if (!it.open) { it.open = 1; it.s0 = it.s1 = i; continue; }
      if (i - it.s1 <= 1) { it.s1 = i; continue; }
      *a = it.s0 + it.seg_start; *b = it.s1 + it.seg_start;
      it.s0 = it.s1 = i;
      return true;
    }
    ++it.level; it.pos = 0; it.open = 0;  // the open run of this level is never emitted
  }
  return false;
}

// areblockscnv on the bin arrays.
RSI_DEVN void blocks_test(const Cta& c, const CandCfg& P, const CandScratch& S, const int* medint, const int* status, int nb, Cnv* list, int nl, Cnv* ov) {
  for (int i = 0; i < nl; ++i) cnv_test(c, P, S, medint, nb, plain_view(list, nl), i, &list[i]);
  Cnv* T = ov;       // entry under test (overlay of list[i])
  Cnv* best = ov + 1;
  for (int i = 0; i < nl; ++i) {
    c.sync();
    const Cnv orig = list[i];
    if (orig.status != -9) continue;
    if ((orig.type == RSIGPU_TYPE_DEL && orig.cnvmed < 0.7 * orig.refmed) || (orig.type == RSIGPU_TYPE_DUP && orig.cnvmed > 1.3 * orig.refmed)) {
      if (c.tid == 0) { list[i].geno = 1; list[i].p1 = P.p; }
      c.sync();
      continue;
    }
    SubsegIter it;
    if (c.tid == 0) { subseg_begin(it, orig, status); *best = orig; }
    int have_pass = 0, bestlen = -1;  // thread 0's view
    for (;;) {
      int a = 0, b = 0, more = 0;
      if (c.tid == 0) more = subseg_next(it, status, &a, &b) ? 1 : 0;
      more = cta_bcast(c, more, 1);
      if (!more) break;
      if (c.tid == 0) { *T = cnv_default(); T->start = a; T->end = b; T->type = orig.type; }
      c.sync();
      ListView V = plain_view(list, nl); V.oi0 = i; V.o0 = T;
      cnv_test(c, P, S, medint, nb, V, i, T);
      if (c.tid == 0 && T->geno != 0) {
        // The reference scans the sub-segments from the last to the first and replaces its pick by
        // every STRICTLY longer passing one (the first passing one unconditionally when the
        // original failed with geno 0) => the longest passing sub-segment, latest-generated on
        // ties; when the original kept geno 1 it additionally has to be longer than the original.
        const bool qualifies = orig.geno == 0 || T->length > orig.length;
        if (qualifies && T->length >= bestlen) { *best = *T; bestlen = T->length; have_pass = 1; }
      }
      c.sync();
    }
    (void)have_pass;
    if (c.tid == 0) list[i] = *best;
    c.sync();
  }
}

// expand_coordinate: compacted index -> reference coordinate through the N-interval table
RSI_DEV int expand_coord(int p, const int* nbeg, const int* nend, int nn) {
  if (nn == 0) return p;
  int dx = 0, prev_brk = 0, prev_inc = 0;
  for (int i = 0; i < nn; ++i) {
    dx += nend[i] - nbeg[i] + 1;
    int brk = nend[i] + 1 - dx;
    if (i == 0 && p < brk) return p;
    if (i > 0 && p >= prev_brk && p < brk) return p + prev_inc;
    prev_brk = brk; prev_inc = dx;
  }
  return p + prev_inc;  // p >= last break
}

// sd_filters (thread 0)
RSI_DEV int sd_filter_list(const CandCfg& P, Cnv* list, int n) {
  const int minlen = P.m * 2 > 500 ? P.m * 2 : 500;
  const double tsd = P.rdsd / 1.2;
  int w = 0;
  for (int j = 0; j < n; ++j) {
    const Cnv& x = list[j];
    bool keep = true;
    int span = x.end - x.start; if (span < 0) span = -span;
    if (span < 1000) keep = false;
    if (x.type == RSIGPU_TYPE_DEL) {
      if (x.p1 > 0.2) keep = false;
      if (x.refsd > 0.6 * tsd) keep = false;
      if (x.cnvsd > 1.3 * tsd) keep = false;
      if (x.cnvsd * P.rdmedian > 2.5 * x.cnvmed * tsd) keep = false;
      const double mn = P.rdmedian < x.refmed ? P.rdmedian : x.refmed;
      if (x.cnvmed < 0.66 * mn && x.cnvsd < tsd && span > 800) keep = true;
    }
    if (x.type == RSIGPU_TYPE_DUP) {
      if (x.p1 > 0.05) keep = false;
      if (x.refsd > 0.6 * tsd) keep = false;
      if (x.cnvsd * P.rdmedian > 2.0 * x.cnvmed * tsd) keep = false;
    }
    if (span < minlen) keep = false;
    if (keep) { if (w != j) list[w] = list[j]; ++w; }
  }
  return w;
}

// Dump buffers for the parity tests (RSIGPU_ARR_SEGMENTS .. DETECTED)
struct CandDumps { Cnv* blocks; int* n_blocks; Cnv* premerge; int* n_premerge; Cnv* merged; int* n_merged; int cap; };
RSI_DEV void dump_list(const Cta& c, const Cnv* list, int n, Cnv* dst, int* ndst, int cap) {
  if (!dst) return;
  for (int j = c.tid; j < n && j < cap; j += c.nthr) dst[j] = list[j];
  if (c.tid == 0) *ndst = n;
  c.sync();
}

// detectcnv from areblockscnv onwards (rsi.cpp:1839-1931), in three stages so that the per-call work
// between them can run one call per thread block:
//   stage A (one block): areblockscnv on the bin arrays, sort, bins -> bases
//   [edge_refine x2 per call: independent calls -> one block each]
//   stage B (one block): sort, mergesegments, sort; scratch offsets for the speculative final tests
//   [final isitcnvwrap per call, speculatively assuming that no earlier call fails: one block each]
//   stage C (one block): accept the speculative results up to the first failing call, redo the rest in
//                        order, score, drop N overlaps, map back to reference coordinates
// `list` holds the rsi segments of one transformation in bin coordinates (nl of them); the result
// (reference coordinates) is left in `list`, its length returned.
RSI_DEVN int cand_stage_a(const Cta& c, const CandCfg& P, const CandScratch& S, int n, const int* medint, const int* status, int nb,
                          Cnv* list, int nl, Cnv* tmp, Cnv* ov, const CandDumps& D, int skip_blocks) {
  long long tm = cand_clock();
  if (!skip_blocks) blocks_test(c, P, S, medint, status, nb, list, nl, ov);
  cand_tick(c, S, 12, &tm);
  dump_list(c, list, nl, D.blocks, D.n_blocks, D.cap);
  list_sort(c, list, nl, tmp);
  c.sync();
  if (c.tid == 0) {   // bins -> bases
    int w = 0;
    for (int j = 0; j < nl; ++j) {
      Cnv x = list[j];
      if (x.geno == 0) continue;
      if (x.start == x.end) continue;
      x.start = x.start * P.m + P.m / 2;
      x.end = x.end * P.m + P.m / 2;
      if (x.start < 0) x.start = 0;
      if (x.end > n - 1) x.end = n - 1;
      x.length = x.end - x.start + 1;
      x.tid = P.tid;
      list[w++] = x;
    }
    nl = w;
  }
  nl = cta_bcast(c, nl, 2);
  cand_tick(c, S, 13, &tm);
  return nl;
}

// number of scratch entries isitcnvwrap may gather for a call on the per-base array (capacity of RDref, rsi.cpp:199)
RSI_DEV long long cand_ref_need(const CandCfg& P, const Cnv& x) {
  int d = x.end - x.start + 1;
  if (d < P.m * P.minmlen) d = (int)(P.m * P.minmlen);
  return (long long)(P.chklen * d * 2) + 64;
}

RSI_DEVN int cand_stage_b(const Cta& c, const CandCfg& P, const CandScratch& S, const int* RD, int n, Cnv* list, int nl, Cnv* tmp, Cnv* ov,
                          const CandDumps& D, long long* spec_off, long long spec_cap, int* spec_on) {
  long long tm = cand_clock();
  list_sort(c, list, nl, tmp);
  dump_list(c, list, nl, D.premerge, D.n_premerge, D.cap);
  nl = merge_segments(c, P, S, RD, n, list, nl, ov);
  cand_tick(c, S, 10, &tm);
  list_sort(c, list, nl, tmp);
  dump_list(c, list, nl, D.merged, D.n_merged, D.cap);
  if (spec_off && c.tid == 0) {   // disjoint scratch slices for the speculative per-call tests
    long long off = 0;
    for (int j = 0; j < nl; ++j) { spec_off[j] = off; off += cand_ref_need(P, list[j]); }
    spec_off[nl] = off;
    *spec_on = off <= spec_cap ? 1 : 0;
  }
  c.sync();
  return nl;
}

// one call's final test into res[j] (the list itself stays as it was, so every block sees the same neighbours)
RSI_DEVN void cand_final_one(const Cta& c, const CandCfg& P, const CandScratch& S, const int* RD, int n, const Cnv* list, int nl, int j, Cnv* res) {
  if (c.tid == 0) res[j] = list[j];
  c.sync();
  ListView V = plain_view(list, nl); V.oi0 = j; V.o0 = &res[j];
  cnv_test(c, P, S, RD, n, V, j, &res[j]);
}

RSI_DEVN int cand_stage_c(const Cta& c, const CandCfg& P, const CandScratch& S, const int* RD, int n, const int* nbeg, const int* nend, int nn,
                          Cnv* list, int nl, const Cnv* res, int spec_valid, int* first_bad_out = nullptr) {
  long long tm = cand_clock();
  // The final tests run in list order and a call that FAILS (status -9) changes the neighbour walk of the
  // calls after it.  The speculative results assumed that nobody failed: they are exact up to and including
  // the first failing call; from there on the tests are redone in order.
  int first_bad = nl;
  if (spec_valid) {
    if (c.tid == 0) { for (int j = 0; j < nl; ++j) if (res[j].status == -9) { first_bad = j; break; } }
    first_bad = cta_bcast(c, first_bad, 5);
    if (first_bad_out && c.tid == 0) *first_bad_out = nl - first_bad;   // number of calls redone in order
    const int upto = first_bad < nl ? first_bad + 1 : nl;
    for (int j = c.tid; j < upto; j += c.nthr) list[j] = res[j];
    c.sync();
  } else first_bad = -1;
  for (int j = (spec_valid ? first_bad + 1 : 0); j < nl; ++j) cnv_test(c, P, S, RD, n, plain_view(list, nl), j, &list[j]);
  if (c.tid == 0) {
    for (int j = 0; j < nl; ++j) {
      Cnv& x = list[j];
      const double len = double(x.end - x.start + 1) / double(P.m);
      x.score = (x.cnvmed - P.rdmedian) * sqrt(len);
      const int p1 = expand_coord(x.start, nbeg, nend, nn), p2 = expand_coord(x.end, nbeg, nend, nn);
      for (int k = 0; k < nn; ++k) { int a = p1 > nbeg[k] ? p1 : nbeg[k], b = p2 < nend[k] ? p2 : nend[k]; if (a <= b) x.status = -9; }
    }
  }
  c.sync();
  cand_tick(c, S, 11, &tm);
  nl = list_drop_deleted(c, list, nl);
  if (c.tid == 0) for (int j = 0; j < nl; ++j) { list[j].start = expand_coord(list[j].start, nbeg, nend, nn); list[j].end = expand_coord(list[j].end, nbeg, nend, nn); }
  c.sync();
  return nl;
}

// the whole stage in one block (tests/hostsim/sim.cpp; also the reference order of operations)
RSI_DEVN int candidates_main(const Cta& c, CandCfg P, const CandScratch& S, const int* RD, int n, const int* medint, const int* status, int nb,
                             const int* nbeg, const int* nend, int nn, Cnv* list, int nl, Cnv* tmp, Cnv* ov, const CandDumps& D, int skip_blocks) {
  nl = cand_stage_a(c, P, S, n, medint, status, nb, list, nl, tmp, ov, D, skip_blocks);
  for (int rep = 0; rep < 2; ++rep) for (int j = 0; j < nl; ++j) edge_refine(c, RD, n, &list[j]);
  nl = cand_stage_b(c, P, S, RD, n, list, nl, tmp, ov, D, nullptr, 0, nullptr);
  return cand_stage_c(c, P, S, RD, n, nbeg, nend, nn, list, nl, nullptr, 0);
}

}  // namespace rsigpu
