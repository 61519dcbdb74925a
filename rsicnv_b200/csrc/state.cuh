// state.cuh -- per-contig device-resident scalars.  The reference keeps these in `rsi::` globals
// (rsi.h:54-122) and passes them between functions on the host; here every kernel reads and writes
// them in HBM so that the whole contig runs as one stream of launches with (almost) no host
// round trips.
#pragma once
#include "rt.cuh"

namespace rsigpu {

enum { GC_WIN = 201, GC_STRATA = 202 };
enum { HIST_ALL_BINS = 65536 };   // value histogram of adjusted depths (apply_cap's median)
enum { MAD_CLASSES = 31 };        // negative_binomial_transfer's strided sub-samples, rsi.cpp:1131
enum { LMAX_CAP = 2048 };         // largest RSI window length the scan kernel is sized for
enum { FQ_BINS_CAP = 1 << 22 };   // buckets of the float (dy = 0.01) histogram quantile
enum { FX_SHIFT = 36 };           // fixed-point scale of bin values in the exact window sums

enum {
  ERR_DEPTH_RANGE = 1 << 0,   // depth < 0 or >= 2^24
  ERR_HIST_RANGE = 1 << 1,    // adjusted depth >= HIST_ALL_BINS
  ERR_LMAX = 1 << 2,          // RSI Lmax > LMAX_CAP
  ERR_FIXEDPOINT = 1 << 3,    // a bin value is not a multiple of 2^-FX_SHIFT or is >= 2^16: window sums would not be exact
  ERR_FQ_BINS = 1 << 4,       // float quantile needs more than FQ_BINS_CAP buckets
  ERR_LISTCAP = 1 << 5,       // more runs / segments than the lists hold
  ERR_CAND = 1 << 6,          // candidate stage scratch overflow (see cand_err)
  ERR_DEGENERATE = 1 << 7,    // MAD == 0 or similar: the reference divides by zero here
  ERR_PILEUP = 1 << 8,        // malformed read batch
  ERR_MAD_RANGE = 1 << 9,
};

struct QuantJob {   // one histogram-quantile evaluation (partition_stat_tp, wufunctions.cpp:363-424)
  u32 omin, omax;   // ordered-uint encodings of min / max
  u32 n;            // number of samples
  u32 pad_;
  double sum;
  double ymin, ymax, dy;
  u64 np;           // buckets
  double q[3];      // lower quartile, median, upper quartile (mean when the range is below dy)
};

// What <out>.log prints of one transformation (rsicnvnbn / rsicnvmed, rsi.cpp:1262-1515): the first-pass values and tables
// are overwritten by the second pass, so they are saved here as they are produced.
struct RsiLogT {
  double tmedian1, tsigma1, tlamda1, target, tmedian2, tsigma2, tlamda2;
  int calmax, lmax, lbreak_del[2], lbreak_dup[2], st_lo, st_hi, leveldel, leveladd, filt_on, used;
  u32 cnt[2][2][LMAX_CAP + 2];                       // [pass][DEL, DUP][L]: bins whose smallest covering window has length L
  float lvl_mean[2 * LMAX_CAP + 3]; u32 lvl_cnt[2 * LMAX_CAP + 3];   // filterstatus' per-level means / counts (index = level - st_lo)
};

struct DevState {
  int err, cand_err;
  // ---- inputs
  int L;                       // contig length
  int Lc;                      // length after N removal (rsi::end - rsi::start + 1)
  int nb;                      // number of bins  = Lc / m
  int m;
  int n_noseq;
  int gc_on, cap_on, trans;
  double cap;                  // -cap
  // ---- pass A: raw depth statistics + GC table (checkgccontent, gccontent.cpp:95-150)
  int rd_min, rd_max;
  u64 pos_sum, pos_cnt;        // sum / count of positive depths
  double rdmean;
  u64 gc_sum[GC_STRATA], gc_cnt[GC_STRATA];
  double gc_tab[GC_STRATA];
  int gstar;                   // #GC in [L-201, L-1]: the recount at the 21st pseudo-slice
  int gc_base, pad_gc_;        // first stratum of pass A's private per-warp tables (set by the host from the contig's GC fraction)
  int s20, r20;                // 20*floor(L/20), L - 20*floor(L/20)
  // ---- cap (apply_cap, loaddata.cpp:229-240)
  double cap_median, cap_thr;
  int capv;
  int hist_base;               // first value of pass B's private histogram window
  // ---- chromosome statistics on the compacted, capped array (rsi.cpp:2202-2203)
  double rdmedian, rdsd, rdmad;
  i64 max_binsum, min_binsum;
  int chist_R;                 // value range of the class histograms (0 .. chist_R-1)
  // ---- negative_binomial_transfer anchors (rsi.cpp:1159-1185); *_raw are written by the host LUT step
  double med_nbt_raw, del_nbt_raw, dup_nbt_raw;
  u32 nb_tmin_ord;             // ordered-uint min of the raw transformed bins
  int pad0_;
  // ---- RSI (rsicnvnbn / rsicnvmed, rsi.cpp:1262-1515)
  double factor;               // sqrt(2(1+eps) ln 3.1e9)
  int Lmax_base;               // max(10000/m, 20)
  int Lmax;
  double tmedian, tsigma, tlamda, target, dev;
  double lim_del, lim_dup;     // 0.75 / 1.25 * RDmedian
  double out_tmedian, out_tlamda;
  u32 cnt_del[LMAX_CAP + 2], cnt_dup[LMAX_CAP + 2];   // #bins whose smallest covering window has length L
  int lbreak_del, lbreak_dup;
  int st_lo, st_hi;            // min / max of the status array
  int last_run_start;          // start of the last run (get_continuous_segments never emits it)
  int n_nonzero;
  int filt_on;                 // filterstatus sanity check passed
  int pad1_;
  double filt_tdel, filt_tadd;
  u32 n_unmarked; u32 pad2_;
  float lvl_sum[2 * LMAX_CAP + 3];   // per-level sequential float sums (index = level - st_lo)
  u32 lvl_cnt[2 * LMAX_CAP + 3];
  // ---- runs / segments
  int n_runs, n_segs;
  // ---- final list
  int n_calls, n_detected;
  int cand_redone, pad3_;      // calls whose speculative final test had to be redone in order
  // ---- pair statistics (bam_rd_pr_stats, pairrd.cpp:112-260)
  int isize_mean, isize_sd;
  // ---- scratch for the quantile jobs
  QuantJob qj[8];
  // ---- for <out>.log: sums of the adjusted / capped depth over all L positions, and the per-transformation record
  double adj_sum, cap_sum; u64 adj_pos;
  RsiLogT rlog[2];             // [0] = NBN, [1] = MED
};

}  // namespace rsigpu
