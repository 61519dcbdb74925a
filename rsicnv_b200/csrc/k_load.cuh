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

// ---------------------------------------------------------------------------------------------
// Geometry of the two per-base passes.  Every WARP is an independent pipeline: it owns the warp-tiles gw, gw + GW, ... (W_T
// consecutive bases each; gw = global warp id), its own two-stage ring in shared memory that bulk copies (cp.async.bulk,
// rt.cuh) fill with the tile's depth words and FASTA bytes (incl. the window halo), its own mbarriers and its own private
// tables -- no block barrier anywhere in the loop, so a slow warp never stalls the other warps of the SM.  Lane l owns the
// W_CH consecutive bases [l*W_CH, (l+1)*W_CH) of the tile: the GC count of its 201-base window moves by at most one per
// base, so it is one popcount of the window at the chunk start plus prefix popcounts of the in/out bits.  W_CH is 4 * odd:
// the 16-byte shared-memory loads of the 8 lanes of a quarter-warp hit 8 different bank groups.
enum { W_CH = 28, W_T = 32 * W_CH /* 896 bases */,
       W_FL = 208 /* FASTA bytes staged in front of the tile: the window of the contig's last bases starts up to 201 before them */,
       W_FR = 112 /* ... and behind it: the window reaches 100 past a base; both multiples of 16 */,
       W_FA = W_T + W_FL + W_FR /* 1216 FASTA bytes per tile */, W_GW = W_FA / 32 /* 38 GC bit words */,
       W_STAGE = W_T * 4 + W_FA /* 4800 bytes */, W_BITS = (W_GW + 2) * 4 };
enum { LD_T = 8 * W_T, LD_TILE = LD_T };   // depth arrays are padded to a multiple of this (rsigpu.cu)
enum { A_NW = 8, A_NT = A_NW * 32, A_ROWS = 64 /* GC strata in a warp's private table */ };
enum { B_NW = 10, B_NT = B_NW * 32, B_K = 128 /* value window of pass B's private histogram */, B_NCACHE = 256 /* N intervals cached in shared memory */ };
static_assert(W_FA % 32 == 0 && (W_CH % 8) == 4 && W_STAGE % 16 == 0, "tile geometry");

// ---------------------------------------------------------------------------------------------
// N runs of the contig (uppercase 'N' only): run starts and run ends are appended (unordered) to two
// lists; the host sorts the few hundred entries, pads and merges them (get_noseq_regions,
// loaddata.cpp:243-273).  Also counts the uppercase G/C bases: the mean GC count of a 201-base window
// centres the stratum window of pass A's private tables.
__global__ void k_n_runs(const u8* __restrict__ fa, int L, int* beg, int* end, int* n_beg, int* n_end, int cap, u64* n_gc) {
  u32 gc = 0;
  for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < L; i += (int)(gridDim.x * blockDim.x)) {
    const u8 ch = fa[i];
    gc += ((ch | 4) == 'G') ? 1u : 0u;          // 'C' | 4 == 'G'
    if (ch != 'N') continue;
    if (i == 0 || fa[i - 1] != 'N') { int k = atomicAdd(n_beg, 1); if (k < cap) beg[k] = i; }
    if (i == L - 1 || fa[i + 1] != 'N') { int k = atomicAdd(n_end, 1); if (k < cap) end[k] = i; }
  }
  gc = __reduce_add_sync(0xffffffffu, gc);
  if ((threadIdx.x & 31) == 0 && gc) atomicAdd(n_gc, (u64)gc);
}

// ---------------------------------------------------------------------------------------------
// The 201-bp window of base i is [lo, lo+200] with lo = clamp(i-100, 0, L-202) (gccontent.cpp:124-132 incl. the "right
// edge never takes the last base" quirk).
__device__ __forceinline__ int gc_lo(int i, int L) { return iclamp(i - GC_WIN / 2, 0, L - GC_WIN - 1); }

// one lane: arm the stage's barrier and issue the copies of warp-tile `wt` (depth words; FASTA bytes [w0 - W_FL, w0 + W_T + W_FR))
__device__ __forceinline__ void w_issue(unsigned char* stage, u64* bar, int wt, const int* __restrict__ rd, const u8* __restrict__ fa, int do_gc) {
  const size_t w0 = (size_t)wt * W_T;
  const u32 fskip = wt == 0 ? W_FL : 0;     // nothing in front of base 0
  const u32 fbytes = do_gc ? W_FA - fskip : 0;
  fence_proxy_async();
  mbar_expect_tx(bar, (u32)W_T * 4 + fbytes);
  bulk_g2s(stage, rd + w0, (u32)W_T * 4, bar);
  if (fbytes) bulk_g2s(stage + (size_t)W_T * 4 + fskip, fa + w0 - W_FL + fskip, fbytes, bar);
}
// GC flags of the staged FASTA bytes as a bit string: bit k of gcb <-> base w0 - W_FL + k.  One word per lane (32 bytes), twice.
__device__ __forceinline__ void w_gc_bits(const unsigned char* stage_fa, u32* gcb, int wt, int lane) {
  const uint4* f = reinterpret_cast<const uint4*>(stage_fa);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int w = r * 32 + lane;
    if (w < W_GW) {
      u32 word = 0;
      if (!(wt == 0 && w * 32 < W_FL)) {      // bytes in front of base 0 were never copied (and are never used)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 v = f[w * 2 + h];
          const u32 x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const u32 y = (x[k] | 0x04040404u) ^ 0x47474747u;                      // zero byte <=> 'G' or 'C' (uppercase only)
            const u32 z = ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);   // 0x80 in every zero byte
            word |= ((((z >> 7) * 0x00204081u) >> 21) & 0xfu) << (h * 16 + k * 4);
          }
        }
      }
      gcb[w] = word;
    }
  }
  if (lane < 2) gcb[W_GW + lane] = 0;
}
// number of G/C among the 201 bases that start at bit a of the tile's bit string
__device__ __forceinline__ int w_gc_count(const u32* gcb, int a) {
  const int w0 = a >> 5, sh = a & 31;
  int n = 0;
  u32 cur = gcb[w0];
#pragma unroll
  for (int k = 0; k < 7; ++k) {          // 6 whole words + 9 bits
    const u32 nxt = gcb[w0 + k + 1];
    const u32 v = (u32)((((u64)nxt << 32) | cur) >> sh);
    n += __popc(k < 6 ? v : (v & 0x1ffu));
    cur = nxt;
  }
  return n;
}
__device__ __forceinline__ u32 w_bits32(const u32* gcb, int a) {
  const int w0 = a >> 5;
  return (u32)((((u64)gcb[w0 + 1] << 32) | gcb[w0]) >> (a & 31));
}
// is every base of the warp-tile an interior position of the contig (window unclamped, whole tile inside the contig)?
__device__ __forceinline__ bool w_interior(int w0, int L) { return w0 >= GC_WIN / 2 && w0 + W_T - 1 <= L - GC_WIN / 2 - 2; }

// ---------------------------------------------------------------------------------------------
// Pass A: mean of the positive depths and the per-stratum depth sums / counts (5 B/base read; checkgccontent's table,
// gccontent.cpp:95-150).  Every warp has a PRIVATE table wtab[A_ROWS + 1][32 lanes] of packed (count << 40 | sum) words: lane
// l only ever touches column l, so the per-base update is a plain conflict-free 8-byte read-modify-write -- no atomics, no
// warp votes.  Rows cover A_ROWS = 64 of the 202 strata, and the window FOLLOWS the data: a warp works through a contiguous
// range of warp-tiles (the GC level of a genome changes over tens of kilobases, far slower than a tile), checks the mean
// window count of every tile and, when it has drifted out of the middle half of the rows, adds its table into the block
// table and re-centres (a few hundred instructions against ~7 k per tile).  A base outside the rows goes to the spare row
// (never read) and is redone through the block table with shared-memory atomics -- rare now; with ONE window per contig,
// centred on the contig's mean, 40 % of the kernel's stall samples sat on those atomics (profiles/r2_streaming_gc_table_lines.txt).  The update is
// branch-free and software-pipelined: the next base's word is loaded BEFORE the current one is stored, and forwarded in
// registers when both are the same row (every second base), so a lane's chain of updates never waits for shared memory.
// A lane adds < 2^16 bases of depth < 2^24 to one word.
// Dynamic shared memory per warp: 2 stages | bit string | wtab;  per block: osum[GC_STRATA] u64 | ocnt[GC_STRATA] u32
#define RSI_A_WARP_BYTES ((size_t)2 * W_STAGE + W_BITS + (size_t)(A_ROWS + 1) * 256)
#define RSI_SMEM_A ((size_t)A_NW * RSI_A_WARP_BYTES + (size_t)GC_STRATA * 12 + 16)
// wt0 .. wt1: the warp-tiles this launch covers (the whole contig, or one part of a contig split over several GPUs)
// the warp's private table [A_ROWS + 1][32 lanes] added into the block table at strata gbase .. gbase + A_ROWS - 1 and cleared
// (row A_ROWS only parks the bases whose stratum lies outside: they are counted on the overflow path)
__device__ __forceinline__ void a_flush_rows(u64* wtab, int lane, int gbase, u64* osum, u32* ocnt) {
  __syncwarp();
  for (int row = lane; row < A_ROWS; row += 32) {
    u64 sum = 0, cnt = 0;
    for (int l = 0; l < 32; ++l) { u64* e = &wtab[(size_t)row * 32 + ((l + lane) & 31)]; sum += *e & ((1ull << 40) - 1); cnt += *e >> 40; *e = 0ull; }
    const int g = gbase + row;
    if (cnt && g < GC_STRATA) { atomicAdd(&osum[g], sum); atomicAdd(&ocnt[g], (u32)cnt); }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(A_NT) k_gc_table(const int* __restrict__ rd, const u8* __restrict__ fa, DevState* st, int wt0, int wt1) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  __shared__ __align__(8) u64 s_bar[A_NW * 2];
  const int L = st->L, do_gc = st->gc_on;
  int gbase = st->gc_base;             // first stratum of THIS warp's private table: re-centred when the local GC level drifts away
  const int tid = c.tid, lane = tid & 31, warp = tid >> 5;
  unsigned char* wsm = smem + (size_t)warp * RSI_A_WARP_BYTES;
  u32* gcb = reinterpret_cast<u32*>(wsm + 2 * W_STAGE);
  u64* wtab = reinterpret_cast<u64*>(wsm + 2 * W_STAGE + W_BITS);
  u64* osum = reinterpret_cast<u64*>(smem + (size_t)A_NW * RSI_A_WARP_BYTES);
  u32* ocnt = reinterpret_cast<u32*>(osum + GC_STRATA);
  u64* bar = s_bar + warp * 2;
  if (do_gc) {
    for (int k = lane; k < (A_ROWS + 1) * 32; k += 32) wtab[k] = 0ull;
    for (int k = tid; k < GC_STRATA; k += A_NT) { osum[k] = 0ull; ocnt[k] = 0u; }
  }
  if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  c.sync();
  u64* col = wtab + lane;
  const int nwt = imin((L + W_T - 1) / W_T, wt1);
  // every warp takes a CONTIGUOUS range of warp-tiles (~50 kb of a chr19-sized contig): the GC level of neighbouring
  // tiles is similar, so the 64 rows of the private table can follow it
  const int GW = (int)gridDim.x * A_NW, per = (imax(nwt - wt0, 0) + GW - 1) / GW;
  const int gw = wt0 + ((int)blockIdx.x * A_NW + warp) * per, gw_end = imin(gw + per, nwt);
  if (lane == 0)
    for (int s = 0; s < 2; ++s) { const int t = gw + s; if (t < gw_end) w_issue(wsm + s * W_STAGE, &bar[s], t, rd, fa, do_gc); }
  u64 psum = 0, pcnt = 0;
  u64 zsum = 0; u32 zcnt = 0;          // stratum 0 (windows without any G/C: the N stretches) is kept in registers
  int vmin = 0x7fffffff, vmax = -0x7fffffff - 1;
  int it = 0;
  for (int wt = gw; wt < gw_end; ++wt, ++it) {
    const int s = it & 1;
    unsigned char* stage = wsm + s * W_STAGE;
    mbar_wait(&bar[s], (u32)(it >> 1) & 1u);
    const int w0 = wt * W_T;
    if (do_gc) { w_gc_bits(stage + (size_t)W_T * 4, gcb, wt, lane); __syncwarp(); }
    const int* x4 = reinterpret_cast<const int*>(stage) + lane * W_CH;
    const int p0 = w0 + lane * W_CH;
    u32 tsum = 0, tcnt = 0;
    if (w_interior(w0, L) || (!do_gc && w0 + W_T <= L)) {
      // ---- fast path: 28 interior bases, straight-line code
      int xs[W_CH];
#pragma unroll
      for (int j4 = 0; j4 < W_CH / 4; ++j4) {
        const int4 v = *reinterpret_cast<const int4*>(x4 + j4 * 4);
        xs[j4 * 4] = v.x; xs[j4 * 4 + 1] = v.y; xs[j4 * 4 + 2] = v.z; xs[j4 * 4 + 3] = v.w;
      }
      // depth statistics: a depth outside [0, 2^24) rejects the contig, so the OR of the words is enough as a range check
      // (any negative word sets the sign bit, any word >= 2^24 a bit above 23) and the sums may assume x >= 0
      int orx = 0;
#pragma unroll
      for (int j = 0; j < W_CH; ++j) { const int x = xs[j]; orx |= x; tsum += (u32)x; tcnt += (u32)imin(x, 1); }
      vmin = imin(vmin, orx < 0 ? -1 : 0); vmax = imax(vmax, orx < 0 ? 0x7fffffff : orx);
      if (do_gc) {
        const int q = lane * W_CH;               // p - w0; bit index of base p is q + W_FL
        const int gabs = w_gc_count(gcb, q + W_FL - GC_WIN / 2);
        {   // keep the table centred on the tile's GC level (mean of the 32 chunk-start windows)
          int gm = gabs;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) gm += __shfl_xor_sync(0xffffffffu, gm, o);
          gm >>= 5;
          if (gm != 0 && (gm < gbase + A_ROWS / 4 || gm >= gbase + 3 * A_ROWS / 4)) {
            const int nb = imax(0, imin(gm - A_ROWS / 2, (int)GC_STRATA - A_ROWS));
            if (nb != gbase) { a_flush_rows(wtab, lane, gbase, osum, ocnt); gbase = nb; }
          }
        }
        const int g0 = gabs - gbase;
        const u32 inw = w_bits32(gcb, q + W_FL + GC_WIN / 2 + 1), outw = w_bits32(gcb, q + W_FL - GC_WIN / 2);
        const u32 plus = inw & ~outw, minus = outw & ~inw;      // the window count moves by bit(p+101) - bit(p-100) per base
        u32 ovf = 0;
        if (g0 - __popc(minus) >= 0 && g0 + __popc(plus) < A_ROWS) {
          // the whole chunk stays inside the private window (the count moves at most popc(minus) down, popc(plus) up): no clamping
          unsigned re = (unsigned)g0;
          u64 e = col[re * 32];
#pragma unroll
          for (int j = 0; j < W_CH; ++j) {
            u64 en = 0; unsigned rn = re;
            if (j + 1 < W_CH) {
              const u32 mk = (2u << j) - 1u;
              rn = (unsigned)(g0 + __popc(plus & mk) - __popc(minus & mk));
              en = col[rn * 32];                     // issued before the store below
            }
            e += (1ull << 40) | (u64)(u32)xs[j];
            col[re * 32] = e;
            e = rn == re ? e : en;
            re = rn;
          }
        } else {
          unsigned row = (unsigned)g0;
          unsigned re = row < (unsigned)A_ROWS ? row : (unsigned)A_ROWS;
          ovf |= row < (unsigned)A_ROWS ? 0u : 1u;
          u64 e = col[re * 32];
#pragma unroll
          for (int j = 0; j < W_CH; ++j) {
            u64 en = 0; unsigned rn = re;
            if (j + 1 < W_CH) {
              const u32 mk = (2u << j) - 1u;
              const unsigned nrow = (unsigned)(g0 + __popc(plus & mk) - __popc(minus & mk));
              rn = nrow < (unsigned)A_ROWS ? nrow : (unsigned)A_ROWS;
              ovf |= nrow < (unsigned)A_ROWS ? 0u : (2u << j);
              en = col[rn * 32];
            }
            e += (1ull << 40) | (u64)((u32)xs[j] & 0xffffffu);
            col[re * 32] = e;
            e = rn == re ? e : en;
            re = rn;
          }
        }
        if (ovf) {                                 // strata outside the private window
          if (gbase > 0 && g0 + gbase == 0 && plus == 0u) {      // a chunk inside an N stretch: every window is empty
            zsum += tsum; zcnt += W_CH;                           // (tsum: depths are >= 0 or the contig is rejected)
          } else {
            for (u32 m = ovf; m; m &= m - 1) {
              const int j = __ffs((int)m) - 1;
              const u32 mk = (1u << j) - 1u;
              const int g = g0 + gbase + __popc(plus & mk) - __popc(minus & mk);
              const u32 xv = (u32)x4[j] & 0xffffffu;     // (from the stage: a dynamic index would push xs[] to local memory)
              if (g == 0) { zsum += xv; zcnt += 1; }
              else { atomicAdd(&osum[g], (u64)xv); atomicAdd(&ocnt[g], 1u); }
            }
          }
        }
      }
    } else {
      // ---- contig ends: clamped windows, partial tile
      for (int j = 0; j < W_CH && p0 + j < L; ++j) {
        const int p = p0 + j, x = x4[j];
        vmin = imin(vmin, x); vmax = imax(vmax, x);
        if (x > 0) { tsum += (u32)x; tcnt += 1; }
        if (do_gc) {
          const int g = w_gc_count(gcb, gc_lo(p, L) - w0 + W_FL);
          const unsigned row = (unsigned)(g - gbase);
          if (row < (unsigned)A_ROWS) col[row * 32] += (1ull << 40) | (u64)((u32)x & 0xffffffu);
          else if (g == 0) { zsum += (u32)x & 0xffffffu; zcnt += 1; }
          else { atomicAdd(&osum[g], (u64)((u32)x & 0xffffffu)); atomicAdd(&ocnt[g], 1u); }
        }
      }
    }
    psum += tsum; pcnt += tcnt;
    __syncwarp();                                  // every lane is done with stage s and with the bit string
    const int t2 = wt + 2;
    if (lane == 0 && t2 < gw_end) w_issue(stage, &bar[s], t2, rd, fa, do_gc);
  }
  if (do_gc) a_flush_rows(wtab, lane, gbase, osum, ocnt);
  c.sync();
  if (do_gc) {
    zsum = c.reduce(zsum, SumOp()); zcnt = c.reduce(zcnt, SumOp());
    if (tid == 0 && zcnt) { atomicAdd(&osum[0], zsum); atomicAdd(&ocnt[0], zcnt); }
    c.sync();
    for (int g = tid; g < GC_STRATA; g += A_NT)
      if (ocnt[g]) { atomicAdd(&st->gc_sum[g], osum[g]); atomicAdd(&st->gc_cnt[g], (u64)ocnt[g]); }
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
// Same warp-private rings and GC counts as pass A.  The adjusted value goes back into the staged tile in place (a lane only
// touches its own chunk) and the warp then writes its tile out with coalesced stores, 16 bytes per lane where the
// destination allows.  The value histogram uses one private column per lane again: vh[B_K][32] u16 per warp (a lane adds
// < 2^16 bases to one counter), updated branch-free; values outside the window are redone after the chunk.
// int(RD * mean / tab + 0.5) exactly as the reference rounds it (gccontent.cpp:80: every operation a correctly rounded double
// operation, no FMA), without paying for a double division per base: the quotient is first taken as a product with the
// table's reciprocal -- within a few ulp of the true quotient -- and the division itself is only done when that estimate
// lands so close to an integer boundary that the few ulp could matter (or is not finite).
__device__ __forceinline__ int gc_adjusted(int x, double mean, double2 te) {
  const double a = __dmul_rn((double)x, mean);
  const double r = __dadd_rn(__dmul_rn(a, te.y), 0.5);
  const int n = (int)r;
  const double f = r - (double)n, d = r * 1e-12;
  if (f > d && f < 1.0 - d) return n;
  return (int)(__dadd_rn(__ddiv_rn(a, te.x), 0.5));
}
// Dynamic shared memory per warp: 2 stages | bit string | vh;  per block: tab[GC_STRATA][8] (f64 value, f64 reciprocal) | N intervals (beg, end, cum)
#define RSI_B_WARP_BYTES ((size_t)2 * W_STAGE + W_BITS + (size_t)(B_K + 1) * 64)
#define RSI_SMEM_B ((size_t)B_NW * RSI_B_WARP_BYTES + (size_t)GC_STRATA * 16 * 8 + (size_t)B_NCACHE * 12 + 16)
__global__ void __launch_bounds__(B_NT) k_gc_adjust(const int* __restrict__ rd, const u8* __restrict__ fa, int* __restrict__ rdc,
                                                     const int* __restrict__ nbeg, const int* __restrict__ nend, const int* __restrict__ ncum,
                                                     u32* hist_all, DevState* st, int wt0, int wt1) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  __shared__ __align__(8) u64 s_bar[B_NW * 2];
  const int L = st->L, do_gc = st->gc_on, nn = st->n_noseq;
  const int tid = c.tid, lane = tid & 31, warp = tid >> 5;
  unsigned char* wsm = smem + (size_t)warp * RSI_B_WARP_BYTES;
  u32* gcb = reinterpret_cast<u32*>(wsm + 2 * W_STAGE);
  u16* vh = reinterpret_cast<u16*>(wsm + 2 * W_STAGE + W_BITS);       // [B_K][32]
  double* tab = reinterpret_cast<double*>(smem + (size_t)B_NW * RSI_B_WARP_BYTES);
  int* nc = reinterpret_cast<int*>(tab + GC_STRATA * 16);               // beg[B_NCACHE] | end[B_NCACHE] | cum[B_NCACHE]
  u64* bar = s_bar + warp * 2;
  for (int k = lane; k < B_K * 16; k += 32) reinterpret_cast<u32*>(vh)[k] = 0u;
  // (table value, its reciprocal) per stratum, 8 copies: the lanes of a quarter-warp read 8 different bank quads
  if (do_gc) for (int k = tid; k < GC_STRATA * 8; k += B_NT) { const double t = st->gc_tab[k >> 3]; tab[2 * k] = t; tab[2 * k + 1] = 1.0 / t; }
  const bool ncached = nn <= B_NCACHE;
  if (ncached) for (int k = tid; k < nn; k += B_NT) { nc[k] = nbeg[k]; nc[B_NCACHE + k] = nend[k]; nc[2 * B_NCACHE + k] = ncum[k]; }
  const int* nb_ = ncached ? nc : nbeg; const int* ne_ = ncached ? nc + B_NCACHE : nend; const int* nm_ = ncached ? nc + 2 * B_NCACHE : ncum;
  if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  c.sync();
  u16* vcol = vh + lane;
  const double2* tcol = reinterpret_cast<const double2*>(tab) + (lane & 7);
  const double mean = st->rdmean;
  const int hb = st->hist_base, s20 = st->s20, r20 = st->r20, gstar = st->gstar;
  const int q0 = s20 + r20 - GC_WIN;   // first overwritten position of the pseudo-slice (r20 >= 2)
  const int removed_all = nn ? nm_[nn - 1] + (ne_[nn - 1] - nb_[nn - 1] + 1) : 0;
  int bad = 0;
  u32 zeros = 0;
  int kn = 0;                          // first N interval that ends at or after the current warp-tile (tiles are visited in increasing order)
  const int nwt = imin((L + W_T - 1) / W_T, wt1);
  const int gw = wt0 + (int)blockIdx.x * B_NW + warp, GW = (int)gridDim.x * B_NW;
  if (lane == 0)
    for (int s = 0; s < 2; ++s) { const int t = gw + s * GW; if (t < nwt) w_issue(wsm + s * W_STAGE, &bar[s], t, rd, fa, do_gc); }
  auto count_slow = [&](int x) {       // value histogram, general form
    const unsigned w = (unsigned)(x - hb);
    if (w < (unsigned)B_K) vcol[w * 32] += 1;
    else if (x == 0) zeros += 1;                       // zeros below the window (N stretches): counted in a register
    else if (x >= HIST_ALL_BINS || x < 0) bad = 1;
    else atomicAdd(&hist_all[x], 1u);
  };
  int it = 0;
  for (int wt = gw; wt < nwt; wt += GW, ++it) {
    const int s = it & 1;
    unsigned char* stage = wsm + s * W_STAGE;
    mbar_wait(&bar[s], (u32)(it >> 1) & 1u);
    const int w0 = wt * W_T, w1 = imin(w0 + W_T, L);
    if (do_gc) { w_gc_bits(stage + (size_t)W_T * 4, gcb, wt, lane); __syncwarp(); }
    int* x4 = reinterpret_cast<int*>(stage) + lane * W_CH;
    const int p0 = w0 + lane * W_CH;
    const bool quirk = r20 >= 2 && w0 + W_T > q0 && w0 < q0 + r20;
    if (w0 + W_T <= L && (!do_gc || (w_interior(w0, L) && w0 + W_T <= s20 && !quirk))) {
      // ---- fast path: straight-line code for 28 bases
      int xs[W_CH];
#pragma unroll
      for (int j4 = 0; j4 < W_CH / 4; ++j4) {
        const int4 v = *reinterpret_cast<const int4*>(x4 + j4 * 4);
        xs[j4 * 4] = v.x; xs[j4 * 4 + 1] = v.y; xs[j4 * 4 + 2] = v.z; xs[j4 * 4 + 3] = v.w;
      }
      if (do_gc) {
        const int q = lane * W_CH;
        const int g0 = w_gc_count(gcb, q + W_FL - GC_WIN / 2);
        const u32 inw = w_bits32(gcb, q + W_FL + GC_WIN / 2 + 1), outw = w_bits32(gcb, q + W_FL - GC_WIN / 2);
        const u32 plus = inw & ~outw, minus = outw & ~inw;
#pragma unroll
        for (int j = 0; j < W_CH; ++j) {
          const u32 mk = (1u << j) - 1u;
          const int g = g0 + __popc(plus & mk) - __popc(minus & mk);
          xs[j] = gc_adjusted(xs[j], mean, tcol[g * 8]);
        }
#pragma unroll
        for (int j4 = 0; j4 < W_CH / 4; ++j4) *reinterpret_cast<int4*>(x4 + j4 * 4) = make_int4(xs[j4 * 4], xs[j4 * 4 + 1], xs[j4 * 4 + 2], xs[j4 * 4 + 3]);
      }
      // value histogram: branch-free pipelined updates of the lane's column (the spare slot B_K * 32 takes what is outside the window)
      u32 ovf = 0;
      unsigned wv = (unsigned)(xs[0] - hb);
      unsigned re = wv < (unsigned)B_K ? wv : (unsigned)B_K;
      ovf |= wv < (unsigned)B_K ? 0u : 1u;
      u32 e = vcol[re * 32];
#pragma unroll
      for (int j = 0; j < W_CH; ++j) {
        u32 en = 0; unsigned rn = re;
        if (j + 1 < W_CH) {
          const unsigned nw_ = (unsigned)(xs[j + 1] - hb);
          rn = nw_ < (unsigned)B_K ? nw_ : (unsigned)B_K;
          ovf |= nw_ < (unsigned)B_K ? 0u : (2u << j);
          en = vcol[rn * 32];
        }
        e += 1;
        vcol[re * 32] = (u16)e;
        e = rn == re ? e : en;
        re = rn;
      }
      if (ovf) {
        for (u32 m = ovf; m; m &= m - 1) {
          const int x = x4[__ffs((int)m) - 1];      // (from the stage, which holds the adjusted values)
          if (x == 0) zeros += 1;
          else if (x >= HIST_ALL_BINS || x < 0) bad = 1;
          else atomicAdd(&hist_all[x], 1u);
        }
      }
    } else {
      for (int j = 0; j < W_CH && p0 + j < L; ++j) {
        const int p = p0 + j;
        int x = x4[j];
        if (do_gc && p < s20) {
          int g;
          if (r20 >= 2 && p >= q0 && p < q0 + r20) { x = rd[p + GC_WIN - r20]; g = gstar; }
          else g = w_gc_count(gcb, gc_lo(p, L) - w0 + W_FL);
          x = gc_adjusted(x, mean, tcol[g * 8]);
          x4[j] = x;
        }
        count_slow(x);
      }
    }
    __syncwarp();
    // ---- N-compacted store of the warp-tile, coalesced
    {
      const int* tile_rd = reinterpret_cast<const int*>(stage);
      while (kn < nn && ne_[kn] < w0) ++kn;
      const bool hasn = kn < nn && nb_[kn] < w1;
      const int shift0 = kn < nn ? nm_[kn] : removed_all;
      const int np = w1 - w0;
      if (!hasn) {
        int* dst = rdc + (w0 - shift0);
        const int a = (int)((4 - ((size_t)(w0 - shift0) & 3)) & 3);      // elements up to the first 16-byte aligned destination
        const int nv = np > a ? (np - a) >> 2 : 0;
        if (lane < a && lane < np) dst[lane] = tile_rd[lane];
        for (int v = lane; v < nv; v += 32) {
          const int q = a + v * 4;
          *reinterpret_cast<int4*>(dst + q) = make_int4(tile_rd[q], tile_rd[q + 1], tile_rd[q + 2], tile_rd[q + 3]);
        }
        for (int q = a + nv * 4 + lane; q < np; q += 32) dst[q] = tile_rd[q];
      } else {
        for (int q = lane; q < np; q += 32) {
          const int p = w0 + q;
          int k = kn, sh = shift0;
          while (k < nn && ne_[k] < p) { sh += ne_[k] - nb_[k] + 1; ++k; }
          if (!(k < nn && p >= nb_[k])) rdc[p - sh] = tile_rd[q];
        }
      }
    }
    __syncwarp();
    const int t2 = wt + 2 * GW;
    if (lane == 0 && t2 < nwt) w_issue(stage, &bar[s], t2, rd, fa, do_gc);
  }
  __syncwarp();
  // flush the warp's private histogram: lane w sums row w over the 32 columns
  for (int w = lane; w < B_K; w += 32) {
    u32 sum = 0;
    for (int l = 0; l < 32; ++l) sum += vh[(size_t)w * 32 + ((l + lane) & 31)];
    if (sum) { if (hb + w < HIST_ALL_BINS) atomicAdd(&hist_all[hb + w], sum); else bad = 1; }
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
  {   // for <out>.log: the mean of the adjusted depth before and after the cap (loaddata.cpp:353-356, 531-536) and of its positive part
    const double medl = (fnz == lnz || pick[1] < 0) ? (double)fnz : (double)pick[1];
    const double thr = medl * st->cap; const int capv = (int)(medl * st->cap), cap_on = st->cap_on;
    u64 s1 = 0, s2 = 0, np = 0;
    for (int v = c.tid; v < (int)HIST_ALL_BINS; v += c.nthr) {
      const u64 h = hist_all[v];
      s1 += h * (u64)v; s2 += h * (u64)((cap_on && (double)v > thr) ? capv : v);
      if (v > 0) np += h;
    }
    s1 = c.reduce(s1, SumOp()); s2 = c.reduce(s2, SumOp()); np = c.reduce(np, SumOp());
    if (c.tid == 0) { st->adj_sum = (double)s1; st->cap_sum = (double)s2; st->adj_pos = np; }
  }
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
