// k_inflate_warp.cuh -- BGZF inflate with one WARP per BGZF block: the form used for chunks of fewer than INFW_MAX_BLOCKS
// blocks (rsigpu_bam_feed picks; larger chunks go to the one-lane-per-block kernel of k_bam.cuh).
//
// Replaces  samtools-0.1.18/bgzf.c:277-313 (inflate_block: raw deflate, -15 window, no CRC check) for the `rsicnv rsi -b` path.
//
// Why two kernels.  Measured on B200 (profiles/r2_inflate_*): the lane-per-block kernel is latency-bound -- a launch
// takes 30-38 ms whether it holds 7 k or 36 k blocks, because every warp walks 32 unrelated streams in lock-step
// (~650 warp instructions per ~7 bytes per lane) and a chunk gives an SM only a handful of such warps.  This kernel is
// bound by instruction issue instead (78 % of the issue slots), so its time is proportional to the bytes decoded:
// 11 ms for a 12 Mbp 30x file, 45 ms for a chr19-sized one.  The two cross at about 28 k blocks (a 45 Mbp contig at 30x).
//
// A deflate stream is a serial chain -- the position of every code depends on the lengths of all codes before it --
// so the unit of parallelism is the BGZF block (<= 64 KiB decoded, ~36 k of them in a chr19-sized BAM).  Giving each
// block to one LANE makes the 32 lanes of a warp walk 32 unrelated streams in lock-step: every step costs the union of
// the paths the lanes take (short code / long code / literal / match), the matches need a warp-wide scatter to be copied
// at all, and the measured cost was ~650 warp instructions per ~7 decoded bytes per lane.  Here a block belongs to a
// WARP and all 32 lanes decode the SAME stream redundantly: control flow is uniform (no divergence, a long code costs
// only when it occurs), the Huffman tables are one per warp in shared memory (every look-up is a broadcast read), the
// compressed words are dealt over the lanes and fetched by shuffle (one coalesced 128-byte load per 32 words, the next
// load always in flight), and the lanes do in parallel what a stream does have in parallel: the bytes of an LZ77 match
// (byte i of a match is src[i mod dist], all independent), the canonical-code construction (counts and symbol slots by
// __match_any_sync, code ranges by a warp scan, direct-table entries one per lane) and stored blocks.  A 2048-thread SM
// holds 32 such warps; the kernel is bound by instruction issue, and its time is proportional to the bytes decoded.
#pragma once
namespace rsigpu {
namespace infw {

#ifndef INFW_LB_BITS
#define INFW_LB_BITS 9
#endif
#ifndef INFW_DB_BITS
#define INFW_DB_BITS 6
#endif
#ifndef INFW_WPC
#define INFW_WPC 4
#endif
enum { INFW_MAX_BLOCKS = 28000 };     // chunks with fewer BGZF blocks than this go to the warp-per-block kernel
enum {
  INFW_LB = INFW_LB_BITS,            // bits of the literal/length code's direct table
  INFW_DB = INFW_DB_BITS,            // bits of the distance code's direct table
  INFW_WARPS = INFW_WPC,
  INFW_NT = 32 * INFW_WPC            // threads per CTA: INFW_WPC warps, one BGZF block each
};
// table entry: bit 31 = a literal (then bits 16-23 are the byte); else bits 8-9 kind, bits 16-30 value (base length, base
// distance); bits 4-7 number of extra bits; bits 0-3 code length (entry 0: the code is longer than the direct table)
enum { INFW_LEN = 1, INFW_EOB = 2, INFW_BAD = 3 };

// One canonical code: direct table for short codes, and for the rest the canonical form -- with v = the next 15 stream
// bits, first bit most significant, the codes of length l are exactly lim[l-1] <= v < lim[l] (lim[l] = left-aligned end of
// the length-l range, non-decreasing in l), so the length is a count of limits <= v and the symbol is sym[base[l] + (v >> (15-l))].
struct InfWarp {
  u32 lt[1 << INFW_LB];             // literal/length direct table
  u32 dt[1 << INFW_DB];             // distance direct table
  u16 llim[16]; short lbase[16]; u16 dlim[16]; short dbase[16];
  u16 lsym[288]; u16 dsym[32];     // symbols in code order
  u16 cnt[16]; u16 off[16];        // while a code is built: codes per length, next symbol slot per length
  u8 lens[320];                    // code lengths of a dynamic block (literal/length then distance)
  u8 cl[128];                      // the code-length code's direct table: symbol << 3 | length
  u8 cll[32];                      // its 19 code lengths
};
struct InfConst { u32 lsym[32]; u32 dsym[32]; };   // base | extra bits << 16 for the length symbols 257.. and the distance symbols

__device__ __forceinline__ u32 inf_entry_lit(const InfConst& K, u32 sym, u32 l) {
  if (sym < 256u) return 0x80000000u | (sym << 16) | l;
  if (sym == 256u) return ((u32)INFW_EOB << 8) | l;
  if (sym > 285u) return ((u32)INFW_BAD << 8) | l;
  const u32 k = K.lsym[sym - 257u];
  return ((k & 0xffffu) << 16) | ((u32)INFW_LEN << 8) | ((k >> 16) << 4) | l;
}
__device__ __forceinline__ u32 inf_entry_dist(const InfConst& K, u32 sym, u32 l) {
  if (sym > 29u) return ((u32)INFW_BAD << 8) | l;
  const u32 k = K.dsym[sym];
  return ((k & 0xffffu) << 16) | ((k >> 16) << 4) | l;
}

// the compressed words of one block, dealt over the lanes: lane i of `cur` holds word 32c + i of the stream, `nxt` the chunk after it
struct WBits {
  const u32* words; u32 nwords;    // the whole compressed chunk as aligned words (reads past it return 0)
  u32 g0;                          // word index of the stream's first (partial) word
  u32 widx;                        // words moved into the bit buffer so far
  u32 cur, nxt;
  u64 buf; int cnt;
};
__device__ __forceinline__ u32 wb_load(const WBits& b, u32 g) { return g < b.nwords ? b.words[g] : 0u; }
__device__ __forceinline__ void wb_init(WBits& b, u32 byte_off, int lane) {
  b.g0 = byte_off >> 2;
  const u32 mis = byte_off & 3u;
  b.cur = wb_load(b, b.g0 + (u32)lane); b.nxt = wb_load(b, b.g0 + 32u + (u32)lane);
  const u32 w0 = __shfl_sync(0xffffffffu, b.cur, 0);
  b.buf = (u64)(w0 >> (8u * mis)); b.cnt = 32 - 8 * (int)mis; b.widx = 1;
}
// at least 33 valid bits afterwards
__device__ __forceinline__ void wb_refill(WBits& b, int lane) {
  if (b.cnt <= 32) {
    const u32 w = __shfl_sync(0xffffffffu, b.cur, (int)(b.widx & 31u));
    b.buf |= (u64)w << b.cnt; b.cnt += 32; b.widx++;
    if ((b.widx & 31u) == 0u) { b.cur = b.nxt; b.nxt = wb_load(b, b.g0 + b.widx + 32u + (u32)lane); }
  }
}
__device__ __forceinline__ u32 wb_take(WBits& b, int n) { const u32 v = (u32)b.buf & ((1u << n) - 1u); b.buf >>= n; b.cnt -= n; return v; }
// consume a code of (e & 15) bits and the (e >> 4 & 15) extra bits behind it (<= 28 bits together, all in the low word); returns the extra bits
__device__ __forceinline__ u32 wb_code_extra(WBits& b, u32 e) {
  const u32 l = e & 15u, ex = (e >> 4) & 15u;
  const u32 v = ((u32)b.buf >> l) & ~(0xffffffffu << ex);
  const int t = (int)(l + ex);
  b.buf >>= t; b.cnt -= t;
  return v;
}
// byte offset (in the chunk) of the next unread byte, when the reader stands on a byte boundary
__device__ __forceinline__ u32 wb_byte_pos(const WBits& b) { return 4u * (b.g0 + b.widx) - (u32)(b.cnt >> 3); }

// Canonical code from n code lengths (shared memory), built by the warp.  Returns < 0 for an over-subscribed set, > 0 for an
// incomplete one, 0 for a complete one (or no codes at all); *nzero = symbols without a code.
template <bool LIT>
__device__ int inf_build(InfWarp& W, const InfConst& K, const u8* lens, int n, int lane, int* nzero) {
  constexpr int FB = LIT ? INFW_LB : INFW_DB;
  u16* lim = LIT ? W.llim : W.dlim; short* base = LIT ? W.lbase : W.dbase; u16* sym = LIT ? W.lsym : W.dsym; u32* tab = LIT ? W.lt : W.dt;
  const u32 lt_mask = (1u << lane) - 1u;
  if (lane < 16) W.cnt[lane] = 0;
  __syncwarp();
  for (int s0 = 0; s0 < n; s0 += 32) {
    const int s = s0 + lane;
    const int l = s < n ? (int)lens[s] : 16;
    const u32 m = __match_any_sync(0xffffffffu, l);
    if (s < n && (m & lt_mask) == 0u) W.cnt[l] = (u16)(W.cnt[l] + __popc(m));     // one lane per distinct length
    __syncwarp();
  }
  const u32 c = (lane >= 1 && lane <= 15) ? (u32)W.cnt[lane] : 0u;
  *nzero = (int)W.cnt[0];
  const u32 sh = (lane >= 1 && lane <= 15) ? c << (15 - lane) : 0u;
  u32 acc = sh, offi = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 ua = __shfl_up_sync(0xffffffffu, acc, o), uo = __shfl_up_sync(0xffffffffu, offi, o);
    if (lane >= o) { acc += ua; offi += uo; }
  }
  const u32 total = __shfl_sync(0xffffffffu, acc, 15);
  if (total > 32768u) return -1;
  if (lane < 16) {
    const u32 first = lane ? (acc - sh) >> (15 - lane) : 0u;      // first code of this length
    lim[lane] = (u16)(lane ? acc : 0u);
    base[lane] = (short)((int)(offi - c) - (int)first);
    W.off[lane] = (u16)(offi - c);
  }
  __syncwarp();
  for (int s0 = 0; s0 < n; s0 += 32) {
    const int s = s0 + lane;
    const int l = s < n ? (int)lens[s] : 16;
    const u32 m = __match_any_sync(0xffffffffu, l);
    const bool on = s < n && l != 0;
    if (on) sym[(int)W.off[l] + __popc(m & lt_mask)] = (u16)s;      // symbols of one length in increasing order
    __syncwarp();
    if (on && (m & lt_mask) == 0u) W.off[l] = (u16)(W.off[l] + __popc(m));
    __syncwarp();
  }
  for (int idx = lane; idx < (1 << FB); idx += 32) {
    const u32 v = (__brev((u32)idx) >> (32 - FB)) << (15 - FB);
    int l = 1;
#pragma unroll
    for (int k = 1; k <= FB; ++k) l += v >= (u32)lim[k] ? 1 : 0;
    u32 e = 0;
    if (l <= FB) { const u32 s = sym[(int)base[l] + (int)(v >> (15 - l))]; e = LIT ? inf_entry_lit(K, s, (u32)l) : inf_entry_dist(K, s, (u32)l); }
    tab[idx] = e;
  }
  __syncwarp();
  return total < 32768u ? 1 : 0;
}
// a code longer than the direct table (or no code at all: INFW_BAD).  Lane k compares with lim[k]: the length is a ballot.
template <bool LIT>
__device__ __forceinline__ u32 inf_long_code(const InfWarp& W, const InfConst& K, u64 buf, int lane) {
  const u16* lim = LIT ? W.llim : W.dlim;
  const u32 v = __brev((u32)buf) >> 17;
  const u32 mine = (lane >= 1 && lane <= 14) ? (u32)lim[lane] : 0xffffffffu;
  const int l = 1 + __popc(__ballot_sync(0xffffffffu, v >= mine));
  if (v >= (u32)lim[l]) return ((u32)INFW_BAD << 8) | 15u;
  const int idx = (int)(LIT ? W.lbase[l] : W.dbase[l]) + (int)(v >> (15 - l));
  return LIT ? inf_entry_lit(K, (u32)W.lsym[idx], (u32)l) : inf_entry_dist(K, (u32)W.dsym[idx], (u32)l);
}

// header of a deflate block.  Returns 0 with *syms = 1 when Huffman-coded symbols follow, *syms = 0 for a stored block
// (copied here), or an error code.
__device__ __noinline__ int inf_block_header(WBits& b, InfWarp& W, const InfConst& K, const u8* comp, u32 end_byte, u8* dst, u32 dst_len, u32* o_io, int* last, int* syms, int lane) {
  wb_refill(b, lane);
  *last = (int)wb_take(b, 1);
  const u32 type = wb_take(b, 2);
  *syms = 0;
  if (type == 3u) return 1;
  if (type == 0u) {   // stored
    wb_take(b, b.cnt & 7);
    wb_refill(b, lane);
    const u32 len = wb_take(b, 16); wb_refill(b, lane); const u32 nlen = wb_take(b, 16);
    if ((len ^ 0xffffu) != nlen) return 2;
    const u32 o = *o_io;
    if (o + len > dst_len) return 3;
    const u32 p = wb_byte_pos(b);
    if (p + len > end_byte) return 4;
    for (u32 i = (u32)lane; i < len; i += 32u) dst[o + i] = comp[p + i];
    *o_io = o + len;
    wb_init(b, p + len, lane);
    return 0;
  }
  int nlen, ndist;
  if (type == 1u) {
    nlen = 288; ndist = 32;           // the fixed code (RFC 1951 3.2.6): 32 five-bit distance codes, of which 30 and 31 never occur
    for (int s = lane; s < 288; s += 32) W.lens[s] = (u8)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
    W.lens[288 + lane] = 5;
    __syncwarp();
  } else {
    wb_refill(b, lane);
    nlen = (int)wb_take(b, 5) + 257; ndist = (int)wb_take(b, 5) + 1;
    const int ncode = (int)wb_take(b, 4) + 4;
    if (nlen > 286 || ndist > 30) return 5;
    if (lane < 19) W.cll[lane] = 0;
    __syncwarp();
    for (int i = 0; i < ncode; ++i) {
      wb_refill(b, lane);
      const u32 v = wb_take(b, 3);
      // order: 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15
      const int pos = i < 3 ? 16 + i : i == 3 ? 0 : (i & 1) ? 8 - ((i - 3) >> 1) : 8 + ((i - 4) >> 1);
      if (lane == 0) W.cll[pos] = (u8)v;
    }
    __syncwarp();
    {   // the code-length code: 19 symbols, codes of <= 7 bits, must be complete
      const u32 lt_mask = (1u << lane) - 1u;
      if (lane < 8) W.cnt[lane] = 0;
      __syncwarp();
      const int l = lane < 19 ? (int)W.cll[lane] : 8;
      const u32 m = __match_any_sync(0xffffffffu, l);
      if (lane < 19 && (m & lt_mask) == 0u) W.cnt[l] = (u16)__popc(m);
      __syncwarp();
      int code = 0, tot = 0;
      for (int k = 1; k < 8; ++k) { const int cn = (int)W.cnt[k]; if (lane == 0) W.off[k] = (u16)code; code = (code + cn) << 1; tot += cn << (7 - k); }
      if (tot != 128) return 6;
      __syncwarp();
      if (lane < 19 && l != 0) {
        const u32 cd = (u32)W.off[l] + (u32)__popc(m & lt_mask);
        const u32 rev = __brev(cd) >> (32 - l);
        const u8 e = (u8)((lane << 3) | l);
        for (u32 k = rev; k < 128u; k += (1u << l)) W.cl[k] = e;
      }
      __syncwarp();
    }
    int idx = 0; u32 prev = 0;
    while (idx < nlen + ndist) {
      wb_refill(b, lane);
      const u32 e = W.cl[(u32)b.buf & 127u];
      const int l = (int)(e & 7u); b.buf >>= l; b.cnt -= l;
      const u32 sym = e >> 3;
      if (sym < 16u) { if (lane == 0) W.lens[idx] = (u8)sym; prev = sym; ++idx; }
      else {
        int rep; u32 v = 0;
        if (sym == 16u) { if (idx == 0) return 8; v = prev; rep = 3 + (int)wb_take(b, 2); }
        else if (sym == 17u) rep = 3 + (int)wb_take(b, 3);
        else rep = 11 + (int)wb_take(b, 7);
        if (idx + rep > nlen + ndist) return 9;
        if (lane < rep) W.lens[idx + lane] = (u8)v;
        if (lane + 32 < rep) W.lens[idx + lane + 32] = (u8)v;
        if (lane + 64 < rep) W.lens[idx + lane + 64] = (u8)v;
        if (lane + 96 < rep) W.lens[idx + lane + 96] = (u8)v;
        if (lane + 128 < rep) W.lens[idx + lane + 128] = (u8)v;
        prev = v; idx += rep;
      }
    }
    __syncwarp();
    if (W.lens[256] == 0) return 10;
  }
  int nz;
  int r = inf_build<true>(W, K, W.lens, nlen, lane, &nz);
  if (r < 0 || (r > 0 && nlen - nz != 1)) return 11;       // incomplete only allowed for a single code
  r = inf_build<false>(W, K, W.lens + nlen, ndist, lane, &nz);
  if (r < 0 || (r > 0 && ndist - nz != 1)) return 12;
  *syms = 1;
  return 0;
}

// inflate: one warp per BGZF block of the chunk
#ifndef INFW_MINB
#define INFW_MINB 12     // <= 40 registers: 48 warps per SM
#endif
__global__ void __launch_bounds__(INFW_NT, INFW_MINB) k_bgzf_inflate_warp(const u8* __restrict__ comp, u32 comp_words, const BgzfBlock* __restrict__ blk, int nblk, u8* __restrict__ U,
                                                         int* __restrict__ err) {
  __shared__ __align__(16) InfWarp sw[INFW_WPC];
  __shared__ InfConst K;
  {
    const u16 lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const u16 lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const u16 dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    const u16 dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    const int t = (int)threadIdx.x;
    if (t < 32) { K.lsym[t] = t < 29 ? (u32)lbase[t] | ((u32)lext[t] << 16) : 0u; K.dsym[t] = t < 30 ? (u32)dbase[t] | ((u32)dext[t] << 16) : 0u; }
  }
  __syncthreads();
  const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
  const int k = (int)blockIdx.x * INFW_WPC + warp;
  if (k >= nblk) return;
  const BgzfBlock B = blk[k];
  if (B.dst_len == 0u) return;
  InfWarp& W = sw[warp];
  WBits b; b.words = reinterpret_cast<const u32*>(comp); b.nwords = comp_words;
  wb_init(b, B.src, lane);
  u8* const dst = U + B.dst;
  const u32 dst_len = B.dst_len, end_byte = B.src + B.src_len;
  u32 o = 0; int rc = 0, last = 0;
  u8* const dstl = dst + lane;
  u8* pa = nullptr; u32 pv = 0;       // the piece of the last match that is loaded but not yet stored
  while (!last && rc == 0) {
    int syms;
    { WBits hb = b; u32 ho = o; int hl = 0; rc = inf_block_header(hb, W, K, comp, end_byte, dst, dst_len, &ho, &hl, &syms, lane); b = hb; o = ho; last = hl; }   // out of line: copies keep the reader in registers
    if (rc || !syms) continue;
#define INFW_PUT_LITERAL(e) { const int l_ = (int)((e) & 15u); if (lane == 0) dst[o] = (u8)((e) >> 16); ++o; b.buf >>= l_; b.cnt -= l_; }
    for (;;) {
      wb_refill(b, lane);              // >= 33 bits: three direct-table codes (<= 27 bits), or one code with its extra bits (<= 20)
      u32 e = W.lt[(u32)b.buf & ((1u << INFW_LB) - 1u)];
      if ((int)e < 0) {                // most symbols of BAM data are literals: up to three per refill
        INFW_PUT_LITERAL(e);
        e = W.lt[(u32)b.buf & ((1u << INFW_LB) - 1u)];
        if ((int)e < 0) {
          INFW_PUT_LITERAL(e);
          e = W.lt[(u32)b.buf & ((1u << INFW_LB) - 1u)];
          if ((int)e < 0) INFW_PUT_LITERAL(e);
        }
        if (o > dst_len) { rc = 3; break; }   // (a few bytes past the block at most, and the whole feed fails)
        continue;                      // whatever followed the literals is looked up again behind a refill
      }
      if ((e & 15u) == 0u) e = inf_long_code<true>(W, K, b.buf, lane);
      if ((int)e < 0) { INFW_PUT_LITERAL(e); if (o > dst_len) { rc = 3; break; } continue; }
      const u32 kind = (e >> 8) & 3u;
      if (kind != (u32)INFW_LEN) {
        { const int l = (int)(e & 15u); b.buf >>= l; b.cnt -= l; }
        if (kind == (u32)INFW_EOB) { if (4u * (b.g0 + b.widx) > end_byte + 16u) rc = 17; }      // ran past the payload and its footer
        else rc = 13;
        break;
      }
      const u32 len = (e >> 16) + wb_code_extra(b, e);
      wb_refill(b, lane);
      u32 d = W.dt[(u32)b.buf & ((1u << INFW_DB) - 1u)];
      if ((d & 15u) == 0u) d = inf_long_code<false>(W, K, b.buf, lane);
      if (d & 0x300u) { rc = 15; break; }
      const u32 dist = (d >> 16) + wb_code_extra(b, d);
      if (dist > o) { rc = 16; break; }
      if (o + len > dst_len) { rc = 3; break; }
      // the match: byte i is from[i mod dist], every byte from data that was complete before this match began.  The last
      // (<= 32-byte) piece is loaded now and stored when the next match begins (or at the end): its L2 round trip runs
      // under the decode of the symbols in between.
      if (pa) { *pa = (u8)pv; pa = nullptr; }
      __syncwarp();                                    // the literals (lane 0) and the previous match are visible to every lane
      u8* const op = dstl + o;                         // this lane's byte of the first piece
      const u8* const from = op - dist;
      if (len <= 32u && dist >= len) {                 // the usual case: one piece, no wrap
        if ((u32)lane < len) { pv = *from; pa = op; }
      } else {
        for (u32 i0 = 0; i0 < len; i0 += 32u) {
          const u32 i = i0 + (u32)lane;
          if (i < len) {
            const u32 v = i < dist ? from[i0] : (from - lane)[i % dist];
            if (i0 + 32u >= len) { pv = v; pa = op + i0; } else op[i0] = (u8)v;
          }
        }
      }
      o += len;
    }
#undef INFW_PUT_LITERAL
  }
  if (pa) *pa = (u8)pv;
  if (rc == 0 && o != dst_len) rc = 18;
  if (rc && lane == 0) { atomicOr(err + 1, (int)BAM_ERR_INFLATE); atomicMax(err + 2, rc); }
}

}  // namespace infw
using infw::k_bgzf_inflate_warp;
}  // namespace rsigpu
