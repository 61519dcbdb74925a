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
// Tile geometry of the two per-base passes.  A block of LD_NT threads owns LD_T consecutive bases per iteration; thread t
// owns the LD_CH consecutive bases [t*LD_CH, (t+1)*LD_CH) of the tile (so the GC count of its window moves by at most one
// per base and is tracked incrementally), which it reads from the staged tile with 16-byte shared-memory loads -- LD_CH is
// 4 * odd, so the 8 lanes of a quarter-warp hit 8 different 16-byte bank groups.  Tiles (depth words + FASTA bytes incl. the
// 100-base window halo on both sides) arrive by bulk copy into a two-stage ring (rt.cuh).
enum { LD_NT = 256, LD_NW = LD_NT / 32, LD_CH = 28, LD_T = LD_NT * LD_CH /* 7168 bases */,
       LD_FPAD = 208 /* FASTA bytes staged in front of the tile: the window of the contig's last bases starts up to 201 before them */,
       LD_FR = 112 /* ... and behind it: the window reaches 100 past a base; both multiples of 16 */,
       LD_FA = LD_T + LD_FPAD + LD_FR /* FASTA bytes per tile */, LD_GW = LD_FA / 32 /* GC bit words */, LD_ROWS = 64 /* GC strata in a warp's private table */ };
enum { B_K = 128 };        // value window of pass B's private histogram
enum { LD_TILE = LD_T };   // depth arrays are padded to a multiple of this (rsigpu.cu)
static_assert(LD_FA % 32 == 0 && (LD_CH % 8) == 4, "tile geometry");

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

struct LdStage { int* rd; u8* fa; };
struct LdRing {   // dynamic shared memory: 2 x (depth tile | FASTA tile) | GC bit words
  LdStage st[2]; u32* gcb; u64* bar;
};
__device__ __forceinline__ LdRing ld_ring(unsigned char* smem, u64* bars) {
  LdRing R;
  R.st[0].rd = reinterpret_cast<int*>(smem); R.st[0].fa = smem + (size_t)LD_T * 4;
  R.st[1].rd = reinterpret_cast<int*>(smem + (size_t)LD_T * 4 + LD_FA); R.st[1].fa = smem + (size_t)LD_T * 8 + LD_FA;
  R.gcb = reinterpret_cast<u32*>(smem + 2 * ((size_t)LD_T * 4 + LD_FA));
  R.bar = bars;
  return R;
}
#define RSI_LD_RING_BYTES (2 * ((size_t)LD_T * 4 + LD_FA) + (size_t)(LD_GW + 4) * 4)
// one thread: arm the stage's barrier and issue the two copies of tile `tile` (depth words; FASTA bytes [t0-LD_FPAD, t0+LD_T+LD_FR))
__device__ __forceinline__ void ld_issue(const LdRing& R, int s, int tile, const int* __restrict__ rd, const u8* __restrict__ fa, int do_gc) {
  const size_t t0 = (size_t)tile * LD_T;
  const u32 fskip = tile == 0 ? LD_FPAD : 0;     // nothing in front of base 0
  const u32 fbytes = do_gc ? LD_FA - fskip : 0;
  fence_proxy_async();
  mbar_expect_tx(&R.bar[s], (u32)LD_T * 4 + fbytes);
  bulk_g2s(R.st[s].rd, rd + t0, (u32)LD_T * 4, &R.bar[s]);
  if (fbytes) bulk_g2s(R.st[s].fa + fskip, fa + t0 - LD_FPAD + fskip, fbytes, &R.bar[s]);
}
// GC flags of the staged FASTA tile as a bit string: bit k of gcb <-> base t0 - LD_FPAD + k.  32 bytes -> one word per thread.
__device__ __forceinline__ void ld_gc_bits(const LdRing& R, int s, int tile, int tid) {
  const uint4* f = reinterpret_cast<const uint4*>(R.st[s].fa);
  for (int w = tid; w < LD_GW; w += LD_NT) {
    u32 word = 0;
    if (!(tile == 0 && w * 32 < LD_FPAD)) {      // bytes in front of base 0 were never copied (and are never used)
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
    R.gcb[w] = word;
  }
  if (tid < 4) R.gcb[LD_GW + tid] = 0;
}
// number of G/C among the 201 bases that start at bit a of the tile's bit string
__device__ __forceinline__ int ld_gc_count(const u32* gcb, int a) {
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
__device__ __forceinline__ u32 ld_bits32(const u32* gcb, int a) {
  const int w0 = a >> 5;
  return (u32)((((u64)gcb[w0 + 1] << 32) | gcb[w0]) >> (a & 31));
}

// ---------------------------------------------------------------------------------------------
// Pass A: mean of the positive depths and the per-stratum depth sums / counts (5 B/base read; checkgccontent's table,
// gccontent.cpp:95-150).  Every warp has a PRIVATE table wtab[LD_ROWS][32 lanes] of packed (count << 40 | sum) words: lane l
// only ever touches column l, so the per-base update is a plain conflict-free 8-byte read-modify-write -- no atomics, no
// warp votes.  Rows cover the strata [gc_base, gc_base + LD_ROWS) around the contig's mean GC count; a base outside that
// window goes to a small block table with shared-memory atomics.  A lane adds < 2^16 bases of depth < 2^24 to one word.
// Dynamic shared memory: ring | wtab[LD_NW][LD_ROWS][32] u64 | osum[GC_STRATA] u64 | ocnt[GC_STRATA] u32
#define RSI_SMEM_A (RSI_LD_RING_BYTES + (size_t)LD_NW * LD_ROWS * 32 * 8 + (size_t)GC_STRATA * 12 + 16)
__global__ void __launch_bounds__(LD_NT) k_gc_table(const int* __restrict__ rd, const u8* __restrict__ fa, DevState* st) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  __shared__ __align__(8) u64 s_bar[2];
  const int L = st->L, do_gc = st->gc_on, gbase = st->gc_base;
  const LdRing R = ld_ring(smem, s_bar);
  u64* wtab_all = reinterpret_cast<u64*>(smem + ((RSI_LD_RING_BYTES + 15) & ~(size_t)15));
  u64* osum = wtab_all + (size_t)LD_NW * LD_ROWS * 32;
  u32* ocnt = reinterpret_cast<u32*>(osum + GC_STRATA);
  const int tid = c.tid, lane = tid & 31, warp = tid >> 5;
  if (do_gc) {
    for (int k = tid; k < LD_NW * LD_ROWS * 32; k += LD_NT) wtab_all[k] = 0ull;
    for (int k = tid; k < GC_STRATA; k += LD_NT) { osum[k] = 0ull; ocnt[k] = 0u; }
  }
  u64* col = wtab_all + (size_t)warp * LD_ROWS * 32 + lane;
  const int ntiles = (L + LD_T - 1) / LD_T;
  if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
  c.sync();
  if (tid == 0)
    for (int s = 0; s < 2; ++s) { const int t = (int)blockIdx.x + s * (int)gridDim.x; if (t < ntiles) ld_issue(R, s, t, rd, fa, do_gc); }
  u64 psum = 0, pcnt = 0;
  int vmin = 0x7fffffff, vmax = -0x7fffffff - 1;
  int it = 0;
  for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x, ++it) {
    const int s = it & 1;
    mbar_wait(&R.bar[s], (u32)(it >> 1) & 1u);
    const int t0 = tile * LD_T;
    if (do_gc) { ld_gc_bits(R, s, tile, tid); c.sync(); }
    const int* x4 = R.st[s].rd + tid * LD_CH;
    const int p0 = t0 + tid * LD_CH;
    u32 tsum = 0, tcnt = 0;
    auto add = [&](int x, int g) {
      vmin = imin(vmin, x); vmax = imax(vmax, x);
      if (x > 0) { tsum += (u32)x; tcnt += 1; }
      if (do_gc) {
        const unsigned row = (unsigned)(g - gbase);
        const u64 e = (1ull << 40) | (u64)((u32)x & 0xffffffu);
        if (row < (unsigned)LD_ROWS) col[row * 32] += e;
        else { atomicAdd(&osum[g], (u64)((u32)x & 0xffffffu)); atomicAdd(&ocnt[g], 1u); }
      }
    };
    if (p0 + LD_CH <= L && (!do_gc || (p0 >= GC_WIN / 2 && p0 + LD_CH - 1 <= L - GC_WIN / 2 - 2))) {
      // interior chunk: the window slides by one base per step, g += bit(p+101) - bit(p-100)
      int g = 0; u32 inw = 0, outw = 0;
      if (do_gc) {
        const int q = tid * LD_CH;               // p - t0; bit index of base p is q + LD_FPAD
        g = ld_gc_count(R.gcb, q + LD_FPAD - GC_WIN / 2);
        inw = ld_bits32(R.gcb, q + LD_FPAD + GC_WIN / 2 + 1); outw = ld_bits32(R.gcb, q + LD_FPAD - GC_WIN / 2);
      }
#pragma unroll
      for (int j4 = 0; j4 < LD_CH / 4; ++j4) {
        const int4 v = *reinterpret_cast<const int4*>(x4 + j4 * 4);
        const int xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j4 * 4 + k;
          add(xs[k], g);
          g += (int)((inw >> j) & 1u) - (int)((outw >> j) & 1u);
        }
      }
    } else {
      for (int j = 0; j < LD_CH && p0 + j < L; ++j) {
        const int p = p0 + j;
        const int g = do_gc ? ld_gc_count(R.gcb, gc_lo(p, L) - t0 + LD_FPAD) : 0;
        add(x4[j], g);
      }
    }
    psum += tsum; pcnt += tcnt;
    c.sync();                                     // everybody is done with stage s (and with the bit string)
    const int t2 = tile + 2 * (int)gridDim.x;
    if (tid == 0 && t2 < ntiles) ld_issue(R, s, t2, rd, fa, do_gc);
  }
  c.sync();
  if (do_gc) {
    // column sums of the private tables: thread (row, part) adds 2 warps x 32 lanes of its row into the block table
    {
      const int row = tid & (LD_ROWS - 1), part = tid / LD_ROWS;          // LD_NT / LD_ROWS = 4 parts
      u64 sum = 0, cnt = 0;
      for (int w = part * (LD_NW / 4); w < (part + 1) * (LD_NW / 4); ++w)
        for (int l = 0; l < 32; ++l) { const u64 e = wtab_all[((size_t)w * LD_ROWS + row) * 32 + ((l + tid) & 31)]; sum += e & ((1ull << 40) - 1); cnt += e >> 40; }
      const int g = gbase + row;
      if (cnt && g < GC_STRATA) { atomicAdd(&osum[g], sum); atomicAdd(&ocnt[g], (u32)cnt); }
    }
    c.sync();
    for (int g = tid; g < GC_STRATA; g += LD_NT)
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
// Same tile ring and incremental GC count as pass A.  The adjusted value goes back into the staged tile in place
// (a thread only touches its own chunk) and the block then writes the tile out with coalesced 16-byte stores.  The value
// histogram uses one private column per lane again: vh[warp][B_K][32] u16 (a lane adds < 2^16 bases to one counter).
// Dynamic shared memory: ring | vh[LD_NW][B_K][32] u16 | tab[GC_STRATA][16] f64
#define RSI_SMEM_B (RSI_LD_RING_BYTES + 16 + (size_t)LD_NW * B_K * 32 * 2 + (size_t)GC_STRATA * 16 * 8)
__global__ void __launch_bounds__(LD_NT) k_gc_adjust(const int* __restrict__ rd, const u8* __restrict__ fa, int* __restrict__ rdc,
                                                      const int* __restrict__ nbeg, const int* __restrict__ nend, const int* __restrict__ ncum,
                                                      u32* hist_all, DevState* st) {
  RSI_DYN_SMEM(smem);
  RSI_CTA_SETUP(c);
  __shared__ __align__(8) u64 s_bar[2];
  __shared__ int s_k0, s_hasn;
  const int L = st->L, do_gc = st->gc_on, nn = st->n_noseq;
  const LdRing R = ld_ring(smem, s_bar);
  u16* vh = reinterpret_cast<u16*>(smem + ((RSI_LD_RING_BYTES + 15) & ~(size_t)15));
  double* tab = reinterpret_cast<double*>(vh + (size_t)LD_NW * B_K * 32);
  const int tid = c.tid, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < LD_NW * B_K * 32 / 2; k += LD_NT) reinterpret_cast<u32*>(vh)[k] = 0u;
  if (do_gc) for (int k = tid; k < GC_STRATA * 16; k += LD_NT) tab[k] = st->gc_tab[k >> 4];   // 16 copies: the lanes of a half-warp read 16 different banks pairs
  u16* vcol = vh + (size_t)warp * B_K * 32 + lane;
  const double* tcol = tab + (lane & 15);
  const double mean = st->rdmean;
  const int hb = st->hist_base, s20 = st->s20, r20 = st->r20, gstar = st->gstar;
  const int q0 = s20 + r20 - GC_WIN;   // first overwritten position of the pseudo-slice (r20 >= 2)
  int bad = 0;
  u32 zeros = 0;
  const int ntiles = (L + LD_T - 1) / LD_T;
  if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
  c.sync();
  if (tid == 0)
    for (int s = 0; s < 2; ++s) { const int t = (int)blockIdx.x + s * (int)gridDim.x; if (t < ntiles) ld_issue(R, s, t, rd, fa, do_gc); }
  int it = 0;
  for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x, ++it) {
    const int s = it & 1;
    mbar_wait(&R.bar[s], (u32)(it >> 1) & 1u);
    const int t0 = tile * LD_T, t1 = imin(t0 + LD_T, L);
    if (tid == 0) {  // first interval that ends at or after t0
      int lo = 0, hi = nn;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (nend[mid] < t0) lo = mid + 1; else hi = mid; }
      s_k0 = lo; s_hasn = (lo < nn && nbeg[lo] < t1) ? 1 : 0;
    }
    if (do_gc) ld_gc_bits(R, s, tile, tid);
    c.sync();
    int* x4 = R.st[s].rd + tid * LD_CH;
    const int p0 = t0 + tid * LD_CH;
    auto count = [&](int x) {     // value histogram over every position (N bases included)
      const unsigned w = (unsigned)(x - hb);
      if (w < (unsigned)B_K) vcol[w * 32] += 1;
      else if (x == 0) zeros += 1;                       // zeros below the window (N stretches): counted in a register
      else if (x >= HIST_ALL_BINS || x < 0) bad = 1;
      else atomicAdd(&hist_all[x], 1u);
    };
    if (!do_gc) {
      for (int j = 0; j < LD_CH && p0 + j < L; ++j) count(x4[j]);
    } else if (p0 + LD_CH <= s20 && p0 >= GC_WIN / 2 && p0 + LD_CH - 1 <= L - GC_WIN / 2 - 2 && !(r20 >= 2 && p0 + LD_CH > q0 && p0 < q0 + r20)) {
      const int q = tid * LD_CH;
      int g = ld_gc_count(R.gcb, q + LD_FPAD - GC_WIN / 2);
      const u32 inw = ld_bits32(R.gcb, q + LD_FPAD + GC_WIN / 2 + 1), outw = ld_bits32(R.gcb, q + LD_FPAD - GC_WIN / 2);
#pragma unroll
      for (int j4 = 0; j4 < LD_CH / 4; ++j4) {
        int4 v = *reinterpret_cast<const int4*>(x4 + j4 * 4);
        int xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j4 * 4 + k;
          xs[k] = (int)(__dadd_rn(__ddiv_rn(__dmul_rn((double)xs[k], mean), tcol[g * 16]), 0.5));
          count(xs[k]);
          g += (int)((inw >> j) & 1u) - (int)((outw >> j) & 1u);
        }
        *reinterpret_cast<int4*>(x4 + j4 * 4) = make_int4(xs[0], xs[1], xs[2], xs[3]);
      }
    } else {
      for (int j = 0; j < LD_CH && p0 + j < L; ++j) {
        const int p = p0 + j;
        int x = x4[j];
        if (p < s20) {
          int g;
          if (r20 >= 2 && p >= q0 && p < q0 + r20) { x = rd[p + GC_WIN - r20]; g = gstar; }
          else g = ld_gc_count(R.gcb, gc_lo(p, L) - t0 + LD_FPAD);
          x = (int)(__dadd_rn(__ddiv_rn(__dmul_rn((double)x, mean), tcol[g * 16]), 0.5));
          x4[j] = x;
        }
        count(x);
      }
    }
    c.sync();
    // N-compacted store of the tile: coalesced, 16 bytes per thread where the tile has no N interval
    {
      const int* tile_rd = R.st[s].rd;
      const int k0 = s_k0, hasn = s_hasn, np = t1 - t0;
      const int shift0 = k0 < nn ? ncum[k0] : (nn ? ncum[nn - 1] + (nend[nn - 1] - nbeg[nn - 1] + 1) : 0);
      if (!hasn) {
        int* dst = rdc + (t0 - shift0);
        const int a = (int)((4 - ((size_t)(t0 - shift0) & 3)) & 3);      // elements up to the first 16-byte aligned destination
        const int nv = np > a ? (np - a) >> 2 : 0;
        if (tid < a && tid < np) dst[tid] = tile_rd[tid];
        for (int v = tid; v < nv; v += LD_NT) {
          const int q = a + v * 4;
          *reinterpret_cast<int4*>(dst + q) = make_int4(tile_rd[q], tile_rd[q + 1], tile_rd[q + 2], tile_rd[q + 3]);
        }
        for (int q = a + nv * 4 + tid; q < np; q += LD_NT) dst[q] = tile_rd[q];
      } else {
        for (int q = tid; q < np; q += LD_NT) {
          const int p = t0 + q;
          int k = k0, sh = shift0;
          while (k < nn && nend[k] < p) { sh += nend[k] - nbeg[k] + 1; ++k; }
          if (!(k < nn && p >= nbeg[k])) rdc[p - sh] = tile_rd[q];
        }
      }
    }
    c.sync();
    const int t2 = tile + 2 * (int)gridDim.x;
    if (tid == 0 && t2 < ntiles) ld_issue(R, s, t2, rd, fa, do_gc);
  }
  c.sync();
  for (int w = tid; w < B_K; w += LD_NT) {
    u32 sum = 0;
    for (int k = 0; k < LD_NW; ++k)
      for (int l = 0; l < 32; ++l) sum += vh[((size_t)k * B_K + w) * 32 + ((l + tid) & 31)];
    if (sum) atomicAdd(&hist_all[hb + w], sum);
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
