// TEST INFRASTRUCTURE -- not part of the product.
//
// C-ABI shim over the UNMODIFIED reference objects (yhwu/rsicnv, compiled in place from
// /root/reference by oracle/Makefile.ref into oracle/_ref/).  It lets the Python tests call the
// reference's own functions (which have external linkage but no header) on in-memory arrays, so the
// CPU restatement in oracle/rsi_oracle.cpp and the CUDA path can be pinned function by function.
// Recipe verified in SURVEY.md Appendix C: rsi.cpp is compiled with -Dmain=rsicnv_main.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the resulting
// oracle/_ref/libref_harness.so.
#include <iostream>
#include <fstream>
#include <sstream>
#include <vector>
#include <string>
#include <cstring>
#include <cmath>
using namespace std;
#include "samfunctions.h"
#include "wu2.h"
#include "rsi.h"
#include "wufunctions.h"
#include "gccontent.h"
#include "loaddata.h"
#include "readref.h"
#include "alglibinterface.h"
#include "bam.h"

// reference functions defined in rsi.cpp / loaddata.cpp without a header (file:line = definition)
void isitcnv(Array<int>& RDref, Array<int>& RDcnv, cnv_st& icnv);                          // rsi.cpp:101
void isitcnvwrap(Array<int>& RD, vector<cnv_st>& cnvlist, int cnvidx);                     // rsi.cpp:175
void get_continuous_segments(Array<int>& RDstatus, int d, vector<cnv_st>& seglist);        // rsi.cpp:291
void multisegments(cnv_st& iseg, Array<int>& status, vector<cnv_st>& mseg);                // rsi.cpp:368
void areblockscnv(Array<int>& RDmedint, Array<int>& status, vector<cnv_st>& seglist);      // rsi.cpp:415
void sortcnvstartposition(vector<cnv_st>& cnvlist);                                        // rsi.cpp:549
string cnv_format1(cnv_st& icnv);                                                          // rsi.cpp:581
void mergesegments(Array<int>& RD, vector<cnv_st>& cnvlist);                               // rsi.cpp:694
void optimize_with_derivative(Array<int>& RD, vector<cnv_st>& cnvlist);                    // rsi.cpp:939
void filterstatus(Array<float>& RDtrans, double dev, Array<int>& status);                  // rsi.cpp:1053
void get_rsi_segments(Array<float>& RDmed, Array<int>& status, double tmedian, vector<cnv_st>&); // :1060
void negative_binomial_transfer(Array<int>& RD, int m, Array<float>& RDt);                 // rsi.cpp:1120
void rsistatus(Array<float>&, Array<int>& RDmedint, double tmedian, double tlamda, int Lmax,
               Array<int>& status);                                                        // rsi.cpp:1191
void rsicnvnbn(Array<float>&, Array<int>& RDmedint, Array<int>& status, vector<cnv_st>&);  // rsi.cpp:1262
void median_transfer(Array<int>& RD, int m, Array<float>& RDt);                            // rsi.cpp:1363
void rsicnvmed(Array<float>&, Array<int>& RDmedint, Array<int>& status, vector<cnv_st>&);  // rsi.cpp:1402
int  expand_coordinate(int p1);                                                            // rsi.cpp:1524
void sd_filters(vector<cnv_st>& cnvlist);                                                  // rsi.cpp:1753
void detectcnv(Array<int>& RD, vector<cnv_st>& cnvlist);                                   // rsi.cpp:1795
void get_noseq_regions(string& FASTA);                                                     // loaddata.cpp:243

extern "C" {

// flat mirror of cnv_st (rsi.h:8-51)
struct ref_cnv {
  int tid, type, geno, status, start, end, length, sc1, sc2, pair;
  double score, p1, p2, cnvmed, cnvsd, cnviqr, refmed, refsd, refiqr, q0;
  int rp, pad_;
};

}  // extern "C"

static void to_flat(const cnv_st& c, ref_cnv& o) {
  o.tid = c.tid; o.type = c.type; o.geno = c.geno; o.status = c.status; o.start = c.start; o.end = c.end;
  o.length = c.length; o.sc1 = c.sc1; o.sc2 = c.sc2; o.pair = c.pair; o.score = c.score; o.p1 = c.p1;
  o.p2 = c.p2; o.cnvmed = c.cnvmed; o.cnvsd = c.cnvsd; o.cnviqr = c.cnviqr; o.refmed = c.refmed;
  o.refsd = c.refsd; o.refiqr = c.refiqr; o.q0 = c.q0; o.rp = c.rp; o.pad_ = 0;
}
static void from_flat(const ref_cnv& c, cnv_st& o) {
  o.tid = c.tid; o.type = c.type; o.geno = c.geno; o.status = c.status; o.start = c.start; o.end = c.end;
  o.length = c.length; o.sc1 = c.sc1; o.sc2 = c.sc2; o.pair = c.pair; o.score = c.score; o.p1 = c.p1;
  o.p2 = c.p2; o.cnvmed = c.cnvmed; o.cnvsd = c.cnvsd; o.cnviqr = c.cnviqr; o.refmed = c.refmed;
  o.refsd = c.refsd; o.refiqr = c.refiqr; o.q0 = c.q0; o.rp = c.rp;
}
static vector<cnv_st> list_in(const ref_cnv* a, int n) {
  vector<cnv_st> v(n);
  for (int i = 0; i < n; ++i) from_flat(a[i], v[i]);
  return v;
}
static int list_out(const vector<cnv_st>& v, ref_cnv* a, int cap) {
  int n = (int)v.size();
  for (int i = 0; i < n && i < cap; ++i) to_flat(v[i], a[i]);
  return n;
}

static std::stringstream g_sink;
static std::streambuf* g_cerr_buf = nullptr;

extern "C" {

// silence rsi::dout (tees to cerr and the unopened rsi::fout) -- the log is not under test
void ref_quiet(int on) {
  if (on && !g_cerr_buf) { g_cerr_buf = std::cerr.rdbuf(g_sink.rdbuf()); }
  if (!on && g_cerr_buf) { std::cerr.rdbuf(g_cerr_buf); g_cerr_buf = nullptr; }
  g_sink.str("");
}

// the tunables main()/get_parameters() would set (rsi.cpp:34-98, 1986-2068)
void ref_set_params(int m, int minq, int min_baseQ, double cap, int gcadjust, const char* trans,
                    int merge, double threshold, double epsilon) {
  rsi::m = m; rsi::minq = minq; rsi::min_baseQ = min_baseQ; rsi::cap = cap; rsi::gcadjust = gcadjust != 0;
  rsi::trans = trans; rsi::merge = merge != 0; rsi::threshold = threshold; rsi::epsilon = epsilon;
  rsi::tid = 0; rsi::target_name.clear(); rsi::target_name.push_back("chr"); rsi::plot = false;
  rsi::chklen = 2.5; rsi::maxchkbp = 100000; rsi::minmlen = 3.01; rsi::buffer = 0.05; rsi::p = 0.05;
}
// the undocumented knobs -reflen / -maxchkbp (rsi.cpp:2024-2026); call after ref_set_params
void ref_set_knobs(double chklen, int maxchkbp) { rsi::chklen = chklen; rsi::maxchkbp = maxchkbp; }
void ref_set_state(double RDmedian, double RDsd, int start, int end, int Lmax, double factor) {
  rsi::RDmedian = RDmedian; rsi::RDsd = RDsd; rsi::start = start; rsi::end = end; rsi::Lmax = Lmax;
  rsi::factor = factor;
}
void ref_get_state(double* out) {
  out[0] = rsi::RDmedian; out[1] = rsi::RDsd; out[2] = rsi::start; out[3] = rsi::end; out[4] = rsi::Lmax;
  out[5] = rsi::factor; out[6] = rsi::nbnmedian; out[7] = rsi::nbnlamda; out[8] = rsi::medmedian;
  out[9] = rsi::medlamda;
}
void ref_set_noncode(const int* beg, const int* end, int n) {
  rsi::noncodelist.clear();
  for (int i = 0; i < n; ++i) { cnv_st c; c.tid = rsi::tid; c.start = beg[i]; c.end = end[i]; rsi::noncodelist.push_back(c); }
}
int ref_get_noncode(int* beg, int* end, int cap) {
  int n = (int)rsi::noncodelist.size();
  for (int i = 0; i < n && i < cap; ++i) { beg[i] = rsi::noncodelist[i].start; end[i] = rsi::noncodelist[i].end; }
  return n;
}

// ---- L0 numeric utilities -------------------------------------------------------------------
double ref_median_i32(int* x, long n) { return _median(x, (size_t)n); }
double ref_median_f32(float* x, long n) { return _median(x, (size_t)n); }
double ref_median_f64(double* x, long n) { return _median(x, (size_t)n); }
double ref_iqr_i32(int* x, long n) { return _interquartilerange(x, (size_t)n); }
double ref_iqr_f32(float* x, long n) { return _interquartilerange(x, (size_t)n); }
double ref_alglib_median_i32(int* x, long n) { return alglib::median(x, (size_t)n); }
double ref_pnorm(double x) { return alglib::pnorm(x); }
double ref_variance_i32(const int* x, int n) {
  Array<int> a(n, x); return variance(a, 0, n - 1, 0.0, -1);
}
double ref_variance_f32(const float* x, int n) {
  Array<float> a(n, x); return variance(a, 0, n - 1, 0.0, -1);
}
void ref_runmean_f32(const float* y, float* smo, int n, int band) {
  Array<float> a(n, y), s(n, 0.0f);
  runmean(a, s, n, band, 1);
  for (int i = 0; i < n; ++i) smo[i] = s[i];
}

// ---- loaders' post-processing (a3,a4,a7,a8,a9) ---------------------------------------------
// get_noseq_regions (loaddata.cpp:243) on an in-memory contig; returns #intervals
int ref_noseq_regions(const char* fasta, int n, int* beg, int* end, int cap) {
  string F(fasta, fasta + n);
  get_noseq_regions(F);
  return ref_get_noncode(beg, end, cap);
}
void ref_checkgccontent(int* rd, const unsigned char* gc, int n) {
  Array<int> RD(n, rd); Array<bool> GC(n);
  for (int i = 0; i < n; ++i) GC[i] = gc[i] != 0;
  checkgccontent(RD, GC);
  for (int i = 0; i < n; ++i) rd[i] = RD[i];
}
double ref_apply_cap(int* rd, int n) {
  Array<int> RD(n, rd);
  apply_cap(RD);
  for (int i = 0; i < n; ++i) rd[i] = RD[i];
  return rsi::RDmedian;
}
int ref_concatenate(int* rd, int n) {
  Array<int> RD(n, rd);
  concatenate_data(RD);
  for (int i = 0; i < RD.size(); ++i) rd[i] = RD[i];
  return RD.size();
}

// ---- transforms (a12,a13) -------------------------------------------------------------------
int ref_median_transfer(const int* rd, int n, int m, float* out) {
  Array<int> RD(n, rd); Array<float> T;
  median_transfer(RD, m, T);
  for (int i = 0; i < T.size(); ++i) out[i] = T[i];
  return T.size();
}
int ref_nb_transfer(const int* rd, int n, int m, float* out) {
  Array<int> RD(n, rd); Array<float> T;
  negative_binomial_transfer(RD, m, T);
  for (int i = 0; i < T.size(); ++i) out[i] = T[i];
  return T.size();
}

// ---- RSI scan (a14-a18) ---------------------------------------------------------------------
void ref_rsistatus(const float* t, const int* medint, int nb, double tmedian, double tlamda, int Lmax,
                   int* status) {
  Array<float> T(nb, t); Array<int> M(nb, medint), S(nb, 0);
  rsistatus(T, M, tmedian, tlamda, Lmax, S);
  for (int i = 0; i < nb; ++i) status[i] = S[i];
}
void ref_filterstatus(const float* t, int nb, double dev, int* status) {
  Array<float> T(nb, t); Array<int> S(nb, status);
  filterstatus(T, dev, S);
  for (int i = 0; i < nb; ++i) status[i] = S[i];
}
int ref_continuous_segments(const int* status, int nb, int d, ref_cnv* out, int cap) {
  Array<int> S(nb, status); vector<cnv_st> v;
  get_continuous_segments(S, d, v);
  return list_out(v, out, cap);
}
int ref_get_rsi_segments(const float* t, const int* status, int nb, double tmedian, ref_cnv* out, int cap) {
  Array<float> T(nb, t); Array<int> S(nb, status); vector<cnv_st> v;
  get_rsi_segments(T, S, tmedian, v);
  return list_out(v, out, cap);
}
// which: 0 = rsicnvnbn, 1 = rsicnvmed
int ref_rsicnv(int which, const float* t, const int* medint, int nb, int* status, ref_cnv* out, int cap) {
  Array<float> T(nb, t); Array<int> M(nb, medint), S(nb, 0); vector<cnv_st> v;
  if (which == 0) rsicnvnbn(T, M, S, v); else rsicnvmed(T, M, S, v);
  for (int i = 0; i < nb; ++i) status[i] = S[i];
  return list_out(v, out, cap);
}

// ---- candidates (a19-a25) -------------------------------------------------------------------
void ref_isitcnvwrap(const int* rd, int n, ref_cnv* list, int nlist, int idx) {
  Array<int> RD(n, rd); vector<cnv_st> v = list_in(list, nlist);
  isitcnvwrap(RD, v, idx);
  list_out(v, list, nlist);
}
int ref_areblockscnv(const int* medint, const int* status, int nb, ref_cnv* list, int nlist) {
  Array<int> M(nb, medint), S(nb, status); vector<cnv_st> v = list_in(list, nlist);
  areblockscnv(M, S, v);
  return list_out(v, list, nlist);
}
void ref_sort(ref_cnv* list, int nlist) {
  vector<cnv_st> v = list_in(list, nlist);
  sortcnvstartposition(v);
  list_out(v, list, nlist);
}
void ref_optimize(const int* rd, int n, ref_cnv* list, int nlist) {
  Array<int> RD(n, rd); vector<cnv_st> v = list_in(list, nlist);
  optimize_with_derivative(RD, v);
  list_out(v, list, nlist);
}
int ref_mergesegments(const int* rd, int n, ref_cnv* list, int nlist) {
  Array<int> RD(n, rd); vector<cnv_st> v = list_in(list, nlist);
  mergesegments(RD, v);
  return list_out(v, list, nlist);
}
int ref_sd_filters(ref_cnv* list, int nlist) {
  vector<cnv_st> v = list_in(list, nlist);
  sd_filters(v);
  return list_out(v, list, nlist);
}
int ref_expand_coordinate(int p) { return expand_coordinate(p); }

// detectcnv on an already compacted depth array (rsi.cpp:1795); globals as set by the caller
int ref_detectcnv(const int* rd, int n, ref_cnv* out, int cap) {
  Array<int> RD(n, rd); vector<cnv_st> v;
  detectcnv(RD, v);
  return list_out(v, out, cap);
}

// The whole depth-file path after text parsing, i.e. what main() does between load_data_from_text's
// parse loop and write_cnv_to_file (loaddata.cpp:481-538, rsi.cpp:2200-2208), on in-memory inputs.
// depth is modified in place to the compacted array; *n_compact receives its length.
// stage: 0 = stop after GC adjust + cap, 1 = stop after concatenate + chr stats, 2 = detectcnv,
//        3 = + sd_filters (the rows that would be written).
int ref_depth_path(int* depth, const char* fasta, int n, int stage, int* n_compact, double* chr_stats,
                   ref_cnv* out, int cap) {
  string F(fasta, fasta + n);
  Array<bool> GC(n);
  for (int k = 0; k < n; ++k) GC[k] = (F[k] == 'G' || F[k] == 'C');
  get_noseq_regions(F);
  Array<int> RD(n, depth);
  rsi::start = 1; rsi::end = RD.size();
  if (rsi::gcadjust) checkgccontent(RD, GC);
  if (rsi::cap > 1) apply_cap(RD);
  *n_compact = n;
  if (stage >= 1) {
    concatenate_data(RD);
    rsi::RDmedian = _median(&RD[0], RD.size());
    rsi::RDsd = sqrt(variance(RD, 0, RD.size() - 1, 0.0, -1));
    *n_compact = RD.size();
  }
  for (int i = 0; i < RD.size(); ++i) depth[i] = RD[i];
  vector<cnv_st> v;
  if (stage >= 2) detectcnv(RD, v);
  if (stage >= 3) sd_filters(v);
  chr_stats[0] = rsi::RDmedian; chr_stats[1] = rsi::RDsd;
  return list_out(v, out, cap);
}

// row formatting (rsi.cpp:581)
int ref_format_row(const ref_cnv* c, const char* chrom, double rdmedian, double rdsd, char* buf, int cap) {
  cnv_st x; from_flat(*c, x);
  rsi::tid = 0; rsi::target_name.clear(); rsi::target_name.push_back(chrom);
  rsi::RDmedian = rdmedian; rsi::RDsd = rdsd;
  string s = cnv_format1(x);
  strncpy(buf, s.c_str(), cap - 1); buf[cap - 1] = 0;
  return (int)s.size();
}

// every alignment record of one refID as the reference's own samtools returns it (bam_read1, bam.c:179-210), in file order:
// the fields the path reads (bam1_core_t, bam1_cigar, bam1_qual).  Returns the number of records, or -1 - needed when a
// capacity is too small; *n_cig / *n_qual = totals.
long ref_bam_records(const char* path, int tid, long cap_reads, long cap_cig, long cap_qual, int* pos, int* mpos, int* isize, int* mtid,
                     unsigned short* flag, unsigned char* mapq, unsigned* cigar_off, unsigned* cigar, unsigned long long* qual_off,
                     unsigned char* qual, long* n_cig, long* n_qual) {
  bamFile fp = bam_open(path, "r");
  if (!fp) return -1;
  bam_header_t* h = bam_header_read(fp);
  bam1_t* b = bam_init1();
  long n = 0, nc = 0, nq = 0;
  bool over = false;
  while (bam_read1(fp, b) >= 0) {
    if (b->core.tid != tid) continue;
    if (n < cap_reads && nc + b->core.n_cigar <= cap_cig && nq + b->core.l_qseq <= cap_qual) {
      pos[n] = b->core.pos; mpos[n] = b->core.mpos; isize[n] = b->core.isize; mtid[n] = b->core.mtid; flag[n] = b->core.flag; mapq[n] = b->core.qual;
      cigar_off[n] = (unsigned)nc; qual_off[n] = (unsigned long long)nq;
      memcpy(cigar + nc, bam1_cigar(b), 4 * (size_t)b->core.n_cigar);
      memcpy(qual + nq, bam1_qual(b), (size_t)b->core.l_qseq);
    } else over = true;
    ++n; nc += b->core.n_cigar; nq += b->core.l_qseq;
  }
  if (!over && n <= cap_reads) { cigar_off[n] = (unsigned)nc; qual_off[n] = (unsigned long long)nq; }
  bam_destroy1(b); bam_header_destroy(h); bam_close(fp);
  *n_cig = nc; *n_qual = nq;
  return over ? -1 - n : n;
}

}  // extern "C"
