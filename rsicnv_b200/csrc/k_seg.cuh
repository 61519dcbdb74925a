// k_seg.cuh -- runs of the final status array and the best-scoring window of every run, then the
// one-block candidate stage.
//
// Replaces (reference file:line relative to src/):
//   get_continuous_segments   rsi.cpp:291-326     k_runs_count, k_runs_scatter  (the last run is never emitted)
//   get_rsi_segments          rsi.cpp:1060-1117   k_run_argmax  (first maximum in (L ascending, start ascending) order)
//   rsicnvnbn/rsicnvmed tail  rsi.cpp:1340-1345   k_candidates, thread 0: keep |score| >= tlamda/2
//   areblockscnv .. detectcnv tail, sd_filters    k_candidates -> candidates.cuh
#pragma once
#include "candidates.cuh"
#include "k_rsi.cuh"

namespace rsigpu {

__global__ void k_runs_count(const int* __restrict__ status, int* __restrict__ tile_rs, DevState* st) {
  RSI_CTA_SETUP(c);
  const int nb = st->nb;
  const int ntiles = (nb + 1023) / 1024;
  for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x) {
    int n = 0;
    for (int j = tile * 1024 + c.tid; j < imin(nb, tile * 1024 + 1024); j += c.nthr) {
      const int s = status[j];
      if (s != 0 && !(j > 0 && same_run(status[j - 1], s))) ++n;
    }
    n = c.reduce(n, SumOp());
    if (c.tid == 0) tile_rs[tile] = n;
  }
}
// runs[2k], runs[2k+1] = first / last bin of run k, in position order, the last run dropped
__global__ void k_runs_scatter(const int* __restrict__ status, const int* __restrict__ tile_rs, int* __restrict__ runs, int cap, DevState* st) {
  RSI_CTA_SETUP(c);
  const int nb = st->nb, last = st->last_run_start;
  const int ntiles = (nb + 1023) / 1024;
  for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x) {
    int off = 0;
    for (int k = c.tid; k < tile; k += c.nthr) off += tile_rs[k];
    off = c.reduce(off, SumOp());
    if (tile == ntiles - 1 && c.tid == 0) {
      int n = off + tile_rs[tile];
      n = n > 0 ? n - 1 : 0;
      if (n > cap) { atomicOr(&st->err, (int)ERR_LISTCAP); n = cap; }
      st->n_runs = n;
    }
    if (tile_rs[tile] == 0) continue;
    for (int j0 = tile * 1024; j0 < imin(nb, tile * 1024 + 1024); j0 += c.nthr) {
      const int j = j0 + c.tid;
      int f = 0;
      if (j < nb && j < tile * 1024 + 1024) { const int s = status[j]; f = (s != 0 && !(j > 0 && same_run(status[j - 1], s))) ? 1 : 0; }
      int tot;
      const int ex = c.scan_excl(f, &tot);
      if (f && j != last && off + ex < cap) {
        int e = j;
        while (e + 1 < nb && same_run(status[e], status[e + 1])) ++e;
        runs[2 * (off + ex)] = j; runs[2 * (off + ex) + 1] = e;
      }
      off += tot;
    }
  }
}

// One block per run: fixed-point prefix of the run in global scratch (pfx[a .. b+1], disjoint per run),
// then every (L, j) window; score = |sum/L - tmedian| * sqrt(L) in double (the reference slides a
// double sum over float bins, exact under the condition k_rsi_scan checks).
struct SegBest { double v; u64 key; };
struct SegBestOp { __device__ __forceinline__ SegBest operator()(SegBest a, SegBest b) const { return (b.v > a.v || (b.v == a.v && b.key < a.key)) ? b : a; } };

__global__ void __launch_bounds__(1024) k_run_argmax(const float* __restrict__ t, const int* __restrict__ status, const int* __restrict__ runs,
                                                     i64* __restrict__ pfx, Cnv* __restrict__ segs, DevState* st) {
  RSI_CTA_SETUP(c);
  const int nruns = st->n_runs;
  const double tmed = st->tmedian;
  for (int r = (int)blockIdx.x; r < nruns; r += (int)gridDim.x) {
    const int a = runs[2 * r], b = runs[2 * r + 1], n = b - a + 1;
    i64* P = pfx + a + r;   // n + 1 entries; "+ r" keeps consecutive runs' (n+1)-entry windows disjoint
    i64 carry = 0;
    c.sync();
    if (c.tid == 0) P[0] = 0;
    for (int j0 = 0; j0 < n; j0 += c.nthr) {
      const int j = j0 + c.tid;
      const i64 v = j < n ? (i64)((double)t[a + j] * 68719476736.0) : 0;
      i64 tot;
      const i64 ex = c.scan_excl_ol(v, &tot);
      if (j < n) P[j + 1] = carry + ex + v;
      carry += tot;
    }
    c.sync();
    // For one length L the score is a monotone function of |sum/L - tmedian| (every rounding step is monotone), so its
    // maximum over the starts is reached at the largest or at the smallest window sum: a thread owns a length and keeps the
    // two extreme sums (integer compares, no double division per window), scores them, and only the winning length is
    // then searched for the FIRST start that reaches the maximum (the reference's order: L ascending, start ascending).
    // Lengths are dealt in mirror pairs (1 + u, n - u): n + 1 windows per pair, and the lanes of a warp have almost equal
    // trip counts; lane u reads P[j + 1 + u]: consecutive addresses.
    SegBest best; best.v = 0.0; best.key = ~0ull;
    const int half = (n + 1) / 2;
    for (int u = c.tid; u < half; u += c.nthr) {
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        const int L = side == 0 ? 1 + u : n - u;
        if (side == 1 && L == 1 + u) break;           // the middle length of an odd n: once
        i64 mx = -0x7fffffffffffffffll - 1, mn = 0x7fffffffffffffffll;
        const i64* PL = P + L;
        for (int j = 0; j + L <= n; ++j) { const i64 sj = PL[j] - P[j]; mx = sj > mx ? sj : mx; mn = sj < mn ? sj : mn; }
        const double dL = (double)L, sL = sqrt(dL);
        const double s_hi = __dmul_rn(fabs(__ddiv_rn((double)mx * (1.0 / 68719476736.0), dL) - tmed), sL);
        const double s_lo = __dmul_rn(fabs(__ddiv_rn((double)mn * (1.0 / 68719476736.0), dL) - tmed), sL);
        const double sc = s_hi > s_lo ? s_hi : s_lo;
        if (sc > best.v || (sc == best.v && (u64)L < best.key)) { best.v = sc; best.key = (u64)L; }
      }
    }
    best = c.reduce_ol(best, SegBestOp());
    if (best.v > 0.0 && best.key != ~0ull) {          // first start of the winning length with that score
      const int L = (int)best.key;
      const double dL = (double)L, sL = sqrt(dL);
      int jf = 0x7fffffff;
      for (int j = c.tid; j + L <= n; j += c.nthr) {
        const double sum = (double)(P[j + L] - P[j]) * (1.0 / 68719476736.0);
        const double sc = __dmul_rn(fabs(__ddiv_rn(sum, dL) - tmed), sL);
        if (sc == best.v) { jf = j; break; }
      }
      jf = c.reduce_ol(jf, MinOp());
      best.key = ((u64)L << 32) | (u64)(unsigned)jf;
    }
    if (c.tid == 0) {
      int ns = a, ne = b;
      if (best.v > 0.0 && best.key != ~0ull) { const int L = (int)(best.key >> 32), j = (int)(best.key & 0xffffffffu); ns = a + j; ne = a + j + L - 1; }
      Cnv x = cnv_default();
      x.start = ns; x.end = ne;
      if (status[ns] > 0) { x.type = RSIGPU_TYPE_DUP; x.score = best.v; } else { x.type = RSIGPU_TYPE_DEL; x.score = -best.v; }
      segs[r] = x;
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct CandArgs {
  const int* rdc; const int* medint; const int* status;
  const int* nbeg; const int* nend;
  Cnv* segs;          // in: one entry per run (k_run_argmax); reused as the working list
  Cnv* tmp; Cnv* ov;  // list-sized scratch, 2-entry overlay scratch
  Cnv* d_segments; Cnv* d_blocks; Cnv* d_premerge; Cnv* d_merged; Cnv* d_detected; Cnv* d_calls;
  int* n_dump;        // [0]=segments [1]=blocks [2]=premerge [3]=merged
  int list_cap;
  CandScratch S;
  int maxchkbp, merge, tid; double chklen;
  int all_phase;      // -ALL (rsi.cpp:1852-1858): 0 = single transformation, 1 = MED half (test and park the segments), 2 = NBN half (test, append, continue)
  Cnv* saved; int* n_saved;
};

enum { CAND_SHIST = 16384 };
struct CandSpec { long long* off; Cnv* res; int* on; int* nl; long long cap; int* ref; long long* pref; float* rm; };

__device__ __forceinline__ CandCfg cand_cfg(const CandArgs& A, const DevState* st) {
  CandCfg P;
  P.m = st->m; P.maxchkbp = A.maxchkbp; P.merge = A.merge; P.tid = A.tid; P.chklen = A.chklen; P.minmlen = 3.01; P.buffer = 0.05; P.p = 0.05;
  P.rdmedian = st->rdmedian; P.rdsd = st->rdsd; P.span = st->Lc;
  return P;
}
__device__ __forceinline__ CandDumps cand_dumps(const CandArgs& A) {
  CandDumps D;
  D.blocks = A.d_blocks; D.n_blocks = &A.n_dump[1]; D.premerge = A.d_premerge; D.n_premerge = &A.n_dump[2];
  D.merged = A.d_merged; D.n_merged = &A.n_dump[3]; D.cap = A.list_cap;
  return D;
}

// stage A: keep |score| >= tlamda/2 (rsi.cpp:1340-1345), areblockscnv, sort, bins -> bases
// The bin-level stage works on a few dozen list entries with many dependent reads by the list-keeping thread:
// the list, its sort buffer and the overlay entries live in shared memory while it runs.
enum { CAND_A_LIST = 192 };
#define RSI_SMEM_CAND_A ((size_t)CAND_SHIST * 4 + (size_t)(2 * CAND_A_LIST + 2) * sizeof(Cnv))
__global__ void __launch_bounds__(1024) k_cand_a(CandArgs A, CandSpec X, DevState* st) {
  RSI_CTA_SETUP(c);
  RSI_DYN_SMEM(smem);
  A.S.shist = reinterpret_cast<unsigned*>(smem); A.S.shist_cap = CAND_SHIST;
  Cnv* slist = reinterpret_cast<Cnv*>(smem + (size_t)CAND_SHIST * 4);
  const CandCfg P = cand_cfg(A, st);
  const int n_runs = st->n_runs;
  int nl = 0;
  if (c.tid == 0) {
    const double half = st->tlamda * 0.5;
    for (int r = 0; r < n_runs; ++r) if (!(fabs(A.segs[r].score) < half)) { if (nl != r) A.segs[nl] = A.segs[r]; ++nl; }
  }
  nl = cta_bcast(c, nl, 3);
  Cnv* const gsegs = A.segs;
  const int nsaved = A.all_phase == 2 ? *A.n_saved : 0;
  const bool in_smem = nl + nsaved <= CAND_A_LIST;
  if (in_smem) {
    for (int j = c.tid; j < nl; j += c.nthr) slist[j] = gsegs[j];
    c.sync();
    A.segs = slist; A.tmp = slist + CAND_A_LIST; A.ov = slist + 2 * CAND_A_LIST;
  }
  dump_list(c, A.segs, nl, A.d_segments, &A.n_dump[0], A.list_cap);
  int skip = 0;
  if (A.all_phase) {
    blocks_test(c, P, A.S, A.medint, A.status, st->nb, A.segs, nl, A.ov);
    c.sync();
    if (A.all_phase == 1) {   // park the tested MED segments; the NBN half follows
      for (int j = c.tid; j < nl; j += c.nthr) A.saved[j] = A.segs[j];
      if (c.tid == 0) *A.n_saved = nl;
      return;
    }
    const int ns = *A.n_saved;   // concatenate: MED list first, then the NBN list (rsi.cpp:1856-1857)
    for (int j = c.tid; j < nl; j += c.nthr) A.tmp[j] = A.segs[j];
    c.sync();
    for (int j = c.tid; j < ns + nl && j < A.list_cap; j += c.nthr) A.segs[j] = j < ns ? A.saved[j] : A.tmp[j - ns];
    c.sync();
    nl = ns + nl < A.list_cap ? ns + nl : A.list_cap;
    skip = 1;
  }
  nl = cand_stage_a(c, P, A.S, st->Lc, A.medint, A.status, st->nb, A.segs, nl, A.tmp, A.ov, cand_dumps(A), skip);
  if (in_smem) { c.sync(); for (int j = c.tid; j < nl; j += c.nthr) gsegs[j] = slist[j]; }
  if (c.tid == 0) *X.nl = nl;
}
// ---- per-call kernels: one thread-block CLUSTER per call.  The CTAs of a cluster act as one big cooperative group
// (Cta with nctas > 1: cluster-wide thread ids, hardware cluster barrier, partial results exchanged through a small
// global scratch), so the neighbourhood of a large call is processed by several SMs instead of one.
#if defined(RSI_SIM)
enum { CAND_CL = 4, CAND_CL_NT = 128 };     // the emulator keeps every fiber of a cluster resident: small shapes
#else
enum { CAND_CL = 8, CAND_CL_NT = 1024 };
#endif
enum { CAND_CL_HIST = 1 << 20 };            // global histogram buckets per cluster
struct ClusterArgs { unsigned char* gx; double* gbc; unsigned* ghist; };
__device__ __forceinline__ Cta cluster_cta(unsigned char* smem, const ClusterArgs& G, int* cid, int* ncl) {
  Cta c;
  const int nct = cluster_size(), rk = cluster_rank();
  c.ltid = (int)threadIdx.x; c.lnthr = (int)blockDim.x; c.nctas = nct; c.rank = rk;
  c.tid = rk * (int)blockDim.x + (int)threadIdx.x; c.nthr = nct * (int)blockDim.x;
  c.red = smem;                                             // 33 * 16 bytes
  *cid = (int)blockIdx.x / nct; *ncl = (int)gridDim.x / nct;
  c.gx = G.gx + (size_t)*cid * CTA_GX_BYTES;
  c.bc = nct > 1 ? G.gbc + (size_t)*cid * 32 : reinterpret_cast<double*>(smem + 33 * 16);
  return c;
}
#define RSI_SMEM_CAND_CL ((size_t)33 * 16 + 32 * 8 + (size_t)CAND_SHIST * 4)

// optimize_with_derivative twice per call; calls are independent of each other
__global__ void __launch_bounds__(1024) k_cand_edge(CandArgs A, CandSpec X, ClusterArgs G, DevState* st) {
  RSI_DYN_SMEM(smem);
  int cid, ncl;
  const Cta c = cluster_cta(smem, G, &cid, &ncl);
  const int nl = *X.nl;
  long long tm = cand_clock();
  for (int j = cid; j < nl; j += ncl) { edge_refine(c, A.rdc, st->Lc, &A.segs[j]); edge_refine(c, A.rdc, st->Lc, &A.segs[j]); }
  if (cid == 0) cand_tick(c, A.S, 9, &tm);
}
__global__ void __launch_bounds__(1024) k_cand_b(CandArgs A, CandSpec X, DevState* st) {
  RSI_CTA_SETUP(c);
  RSI_DYN_SMEM(smem);
  A.S.shist = reinterpret_cast<unsigned*>(smem); A.S.shist_cap = CAND_SHIST;
  const CandCfg P = cand_cfg(A, st);
  int nl = cand_stage_b(c, P, A.S, A.rdc, st->Lc, A.segs, *X.nl, A.tmp, A.ov, cand_dumps(A), X.off, X.cap, X.on);
  c.sync();
  if (c.tid == 0) *X.nl = nl;
}
// speculative final test, one call per cluster, each with its own scratch slice
__global__ void __launch_bounds__(1024) k_cand_final(CandArgs A, CandSpec X, ClusterArgs G, DevState* st) {
  RSI_DYN_SMEM(smem);
  int cid, ncl;
  const Cta c = cluster_cta(smem, G, &cid, &ncl);
  if (!*X.on) return;
  const CandCfg P = cand_cfg(A, st);
  const int nl = *X.nl;
  long long tm = cand_clock();
  for (int j = cid; j < nl; j += ncl) {
    CandScratch S = A.S;
    // one CTA: histogram in its shared memory; several CTAs: the cluster's slice of a global bucket array
    if (c.nctas == 1) { S.shist = reinterpret_cast<unsigned*>(smem + 33 * 16 + 32 * 8); S.shist_cap = CAND_SHIST; S.hist = nullptr; S.hist_cap = 0; }
    else { S.shist = nullptr; S.shist_cap = 0; S.hist = G.ghist + (size_t)cid * CAND_CL_HIST; S.hist_cap = CAND_CL_HIST; }
    S.ref = X.ref + X.off[j]; S.pref = X.pref + X.off[j] + j; S.rm = X.rm + X.off[j];
    S.sub = A.S.sub + (size_t)(1 + cid) * A.S.sub_cap;      // concurrent calls must not share the sub-sampling buffer
    S.ref_cap = (int)(X.off[j + 1] - X.off[j]) - 8;
    S.prof = cid == 0 ? A.S.prof : nullptr;
    cand_final_one(c, P, S, A.rdc, st->Lc, A.segs, nl, j, X.res);
    c.sync();
  }
  if (cid == 0) cand_tick(c, A.S, 6, &tm);
}
__global__ void __launch_bounds__(1024) k_cand_c(CandArgs A, CandSpec X, DevState* st) {
  RSI_CTA_SETUP(c);
  RSI_DYN_SMEM(smem);
  A.S.shist = reinterpret_cast<unsigned*>(smem); A.S.shist_cap = CAND_SHIST;
  const CandCfg P = cand_cfg(A, st);
  int spec = *X.on;
  if (c.tid == 0 && spec && (*A.S.err & (CAND_ERR_HIST | CAND_ERR_REFCAP))) { *A.S.err &= ~(CAND_ERR_HIST | CAND_ERR_REFCAP); spec = 0; }   // a slice overflowed: redo everything in order
  spec = cta_bcast(c, spec, 6);
  int nl = cand_stage_c(c, P, A.S, A.rdc, st->Lc, A.nbeg, A.nend, st->n_noseq, A.segs, *X.nl, X.res, spec, &st->cand_redone);
  for (int j = c.tid; j < nl; j += c.nthr) A.d_detected[j] = A.segs[j];
  c.sync();
  if (c.tid == 0) { st->n_detected = nl; nl = sd_filter_list(P, A.segs, nl); st->n_calls = nl; }
  nl = cta_bcast(c, nl, 4);
  for (int j = c.tid; j < nl; j += c.nthr) A.d_calls[j] = A.segs[j];
  if (c.tid == 0 && *A.S.err) { st->cand_err = *A.S.err; st->err |= ERR_CAND; }
}

}  // namespace rsigpu
