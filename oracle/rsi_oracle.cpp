// TEST INFRASTRUCTURE -- the ORACLE.  Not part of the product; the product (rsicnv_b200/) never
// includes, links or calls anything in this file.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg may load oracle/librsi_oracle.so.
//
// A plain, sequential CPU restatement of the `rsicnv rsi` read-depth -> CNV-call path of
// yhwu/rsicnv, written from the behaviour of the reference (file:line cited per function, relative
// to /root/reference/src).  Parity status: PINNED -- every function below is compared in
// tests/test_oracle_vs_reference.py against the unmodified reference objects (oracle/_ref/
// libref_harness.so, built in place by oracle/Makefile.ref) and against golden vectors produced by
// the reference binary (tests/golden/, generator tests/golden/make_golden.py).  The reference ships
// no tests or fixtures of its own, so those are the only pins that exist.
//
// All arithmetic is done the way the reference does it (x86-64, -O2, no FMA contraction): int32
// depths, fp64 intermediates, float bin arrays, histogram ("partition") quantiles.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

enum { TYPE_DEL = 0, TYPE_DUP = 1, TYPE_UNKNOWN = 2 };

extern "C" struct ocl_cnv {  // flat mirror of cnv_st, rsi.h:8-51 (same layout as ref_cnv / rsigpu_cnv)
  int tid, type, geno, status, start, end, length, sc1, sc2, pair;
  double score, p1, p2, cnvmed, cnvsd, cnviqr, refmed, refsd, refiqr, q0;
  int rp, pad_;
};

ocl_cnv new_cnv() {  // defaults of cnv_st(), rsi.h:30-50
  ocl_cnv c;
  std::memset(&c, 0, sizeof c);
  c.tid = -1; c.type = TYPE_UNKNOWN; c.p1 = 1.0; c.p2 = 1.0; c.q0 = -1.0; c.rp = -1;
  return c;
}

struct Ctx {  // the `rsi::` globals that matter on the path (rsi.h:54-122, defaults rsi.cpp:34-98)
  int m = 101, minq = 0, min_baseQ = 13;
  double cap = 4.0;
  bool gcadjust = true, merge = true;
  int trans = 0;  // 0 NBN, 1 MED, 2 ALL
  double threshold = -1.0, epsilon = 1.5;
  double chklen = 2.5, minmlen = 3.01, buffer = 0.05, p = 0.05;
  int maxchkbp = 100000;
  // per-contig state
  int tid = 0;
  int start = 0, end = 0;  // rsi::start / rsi::end
  int Lmax = -1;
  double factor = 6.6;
  double RDmedian = 0, RDsd = 0;
  double nbnmedian = 0, nbnlamda = 0, medmedian = 0, medlamda = 0;
  std::vector<int> nbeg, nend;  // rsi::noncodelist, 0-based inclusive
};
Ctx G;

// ---------------------------------------------------------------------------------------------
// L0: histogram quantiles -- partition_stat_tp, wufunctions.cpp:363-424 (dy = 1 for int, 0.01 for
// float/double: wufunctions.cpp:425-470).  q[0]=lower quartile, q[1]=median, q[2]=upper quartile.
template <class T>
void hist_stat(const T* x, size_t n, double dy, double q[3]) {
  double ymin = x[0], ymax = x[0], mean = 0;
  for (size_t i = 0; i < n; ++i) {
    mean += x[i];
    if (x[i] < ymin) ymin = x[i];
    if (x[i] > ymax) ymax = x[i];
  }
  mean /= (double)n;
  q[0] = ymin; q[1] = mean; q[2] = ymax;
  if ((ymax - ymin) < dy) return;
  size_t np = (size_t)((ymax - ymin) / dy + 2);
  std::vector<size_t> cnt(np + 1, 0);
  for (size_t i = 0; i < n; ++i) {
    double idx = (x[i] - ymin) / dy + 0.5;
    cnt[(size_t)idx] += 1;
  }
  size_t run = 0, i4 = n / 4, i2 = n / 2, i34 = n * 3 / 4;
  for (size_t b = 0; b < np; ++b) {
    if (run < i4 && run + cnt[b] >= i4) q[0] = ymin + b * dy;
    if (run < i2 && run + cnt[b] >= i2) q[1] = ymin + b * dy;
    if (run < i34 && run + cnt[b] >= i34) q[2] = ymin + b * dy;
    run += cnt[b];
  }
}
double hmedian(const int* x, size_t n) { double q[3]; hist_stat(x, n, 1.0, q); return q[1]; }
double hmedian(const float* x, size_t n) { double q[3]; hist_stat(x, n, 0.01, q); return q[1]; }
double hmedian(const double* x, size_t n) { double q[3]; hist_stat(x, n, 0.01, q); return q[1]; }
double hiqr(const int* x, size_t n) { double q[3]; hist_stat(x, n, 1.0, q); return q[2] - q[0]; }
double hiqr(const float* x, size_t n) { double q[3]; hist_stat(x, n, 0.01, q); return q[2] - q[0]; }

// true sample median -- alglib::median -> samplemedian, alglib/statistics.cpp:3237-3385: element
// (n-1)/2 of the sorted sample, averaged with the next one for even n.
double true_median(const int* x, size_t n) {
  if (n == 0) return 0;
  std::vector<double> v(x, x + n);
  size_t k = (n - 1) / 2;
  std::nth_element(v.begin(), v.begin() + k, v.end());
  double a = v[k];
  if (n % 2 == 1) return a;
  double b = *std::min_element(v.begin() + k + 1, v.end());
  return 0.5 * (a + b);
}

// variance(y,0,n-1,0,-1) -- variancetp end_rule -1, wufunctions.cpp:765-809
template <class T>
double variance_all(const T* y, int n) {
  double s = 0, s2 = 0;
  for (int i = 0; i < n; ++i) { s += (double)y[i]; s2 += (double)y[i] * (double)y[i]; }
  double mean = s / double(n);
  return s2 / double(n) - mean * mean;
}
double mean_range(const int* y, int lo, int hi) {  // mean_tp, wufunctions.cpp:665-689
  double s = 0;
  for (int i = lo; i <= hi; ++i) s += (double)y[i];
  return s / double(hi - lo + 1);
}

// alglib::pnorm -> normaldistribution -> Cephes erf/erfc, alglib/specialfunctions.cpp:3152-3302
double erfc_cephes(double x);
double erf_cephes(double x) {
  double s = x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0);
  x = std::fabs(x);
  if (x < 0.5) {
    double xsq = x * x, p, q;
    p = 0.007547728033418631287834;
    p = -0.288805137207594084924010 + xsq * p;
    p = 14.3383842191748205576712 + xsq * p;
    p = 38.0140318123903008244444 + xsq * p;
    p = 3017.82788536507577809226 + xsq * p;
    p = 7404.07142710151470082064 + xsq * p;
    p = 80437.3630960840172832162 + xsq * p;
    q = 0.0;
    q = 1.00000000000000000000000 + xsq * q;
    q = 38.0190713951939403753468 + xsq * q;
    q = 658.070155459240506326937 + xsq * q;
    q = 6379.60017324428279487120 + xsq * q;
    q = 34216.5257924628539769006 + xsq * q;
    q = 80437.3630960840172826266 + xsq * q;
    return s * 1.1283791670955125738961589031 * x * p / q;
  }
  if (x >= 10) return s;
  return s * (1 - erfc_cephes(x));
}
double erfc_cephes(double x) {
  if (x < 0) return 2 - erfc_cephes(-x);
  if (x < 0.5) return 1.0 - erf_cephes(x);
  if (x >= 10) return 0;
  double p, q;
  p = 0.0;
  p = 0.5641877825507397413087057563 + x * p;
  p = 9.675807882987265400604202961 + x * p;
  p = 77.08161730368428609781633646 + x * p;
  p = 368.5196154710010637133875746 + x * p;
  p = 1143.262070703886173606073338 + x * p;
  p = 2320.439590251635247384768711 + x * p;
  p = 2898.0293292167655611275846 + x * p;
  p = 1826.3348842295112592168999 + x * p;
  q = 1.0;
  q = 17.14980943627607849376131193 + x * q;
  q = 137.1255960500622202878443578 + x * q;
  q = 661.7361207107653469211984771 + x * q;
  q = 2094.384367789539593790281779 + x * q;
  q = 4429.612803883682726711528526 + x * q;
  q = 6089.5424232724435504633068 + x * q;
  q = 4958.82756472114071495438422 + x * q;
  q = 1826.3348842295112595576438 + x * q;
  return std::exp(-(x * x)) * p / q;
}
double pnorm(double x) { return 0.5 * (erf_cephes(x / 1.41421356237309504880) + 1); }

// ---------------------------------------------------------------------------------------------
// a3: N runs (uppercase 'N' only) padded by max(50, m/4), clamped, re-merged.
// get_N_regions readref.cpp:88-112, get_noseq_regions loaddata.cpp:243-273
void noseq_regions(const uint8_t* fasta, int n, int m, std::vector<int>& beg, std::vector<int>& end) {
  std::vector<uint8_t> isn(n);
  for (int i = 0; i < n; ++i) isn[i] = fasta[i] == 'N';
  auto runs = [&](std::vector<int>& b, std::vector<int>& e) {
    b.clear(); e.clear();
    for (int i = 0; i < n; ++i) {
      if (!isn[i]) continue;
      if (!e.empty() && i == e.back() + 1) e.back() = i;
      else { b.push_back(i); e.push_back(i); }
    }
  };
  runs(beg, end);
  int dx = std::max(50, m / 4);
  for (size_t k = 0; k < beg.size(); ++k) {
    int b = std::max(beg[k] - dx, 0), e = std::min(end[k] + dx, n - 1);
    for (int i = b; i <= e; ++i) isn[i] = 1;
  }
  runs(beg, end);
}

// a7: checkgccontent + adjustgccontent, gccontent.cpp:43-184, restated as in SURVEY.md A.3:
// window count of the 201-bp GC window with the "no update when the right edge first touches the
// last base" rule, per-stratum MEAN table, out-of-place map, and the 21st pseudo-slice quirk.
void gc_adjust(int* rd, const uint8_t* gc, int L) {
  const int bin = 201;
  double rdmean = 0; int C = 0;
  for (int i = 0; i < L; ++i) if (rd[i] > 0) { rdmean += rd[i]; ++C; }
  if (C > 0) rdmean /= (double)C;
  std::vector<int> cs(L + 1, 0);
  for (int i = 0; i < L; ++i) cs[i + 1] = cs[i] + (gc[i] ? 1 : 0);
  // sliding rule: the window is [lo, lo+200] with lo = clamp(i-100, 0, L-202)  (gccontent.cpp:124-132)
  auto ngc_slide = [&](int i) { int lo = std::min(std::max(i - bin / 2, 0), L - bin - 1); return cs[lo + bin] - cs[lo]; };
  std::vector<double> tab(bin + 1, 0.0); std::vector<int> cnt(bin + 1, 0);
  for (int i = 0; i < L; ++i) { int g = ngc_slide(i); tab[g] += rd[i]; cnt[g]++; }
  for (int g = 0; g <= bin; ++g) {
    if (cnt[g] > 0) tab[g] /= double(cnt[g]); else tab[g] = rdmean;
    if (tab[g] < 1) tab[g] = rdmean;
  }
  const int S = L / 20, r = L - 20 * S;
  std::vector<int> out(rd, rd + L);
  for (int i = 0; i < 20 * S; ++i) out[i] = (int)(rd[i] * rdmean / tab[ngc_slide(i)] + 0.5);
  if (r >= 2) {  // pseudo-slice: exact recount of [L-201, L-1] at 20S, written 201-r bases too early
    int gstar = cs[L] - cs[L - bin];
    for (int j = 0; j < r; ++j) out[20 * S + r - bin + j] = (int)(rd[20 * S + j] * rdmean / tab[gstar] + 0.5);
  }
  std::memcpy(rd, out.data(), sizeof(int) * (size_t)L);
}

// a8: apply_cap, loaddata.cpp:229-240
double apply_cap(int* rd, int L, double cap) {
  double med = hmedian(rd, (size_t)L);
  for (int i = 0; i < L; ++i) if (rd[i] > med * cap) rd[i] = (int)(med * cap);
  return med;
}

// a9: concatenate_data, loaddata.cpp:48-85
int compact(int* rd, int L, const std::vector<int>& nb, const std::vector<int>& ne) {
  if (nb.empty()) return L;
  std::vector<uint8_t> drop(L, 0);
  for (size_t k = 0; k < nb.size(); ++k) for (int i = nb[k]; i <= ne[k]; ++i) drop[i] = 1;
  int w = 0;
  for (int i = 0; i < L; ++i) if (!drop[i]) rd[w++] = rd[i];
  return w;
}

// a12: median_transfer, rsi.cpp:1363-1379
void median_transfer(const int* rd, int n, int m, std::vector<float>& out) {
  int nb = n / m;
  out.assign(nb, 0.f);
  for (int b = 0; b < nb; ++b) out[b] = (float)true_median(rd + (size_t)b * m, m);
}

// a13: negative_binomial_transfer, rsi.cpp:1120-1188
double nb_formula(double sum, double m2, double r) {
  return 2.0 * std::sqrt(r) * std::log(std::sqrt((sum + 0.25) / (m2 * r - 0.5)) + std::sqrt(1.0 + (sum + 0.25) / (m2 * r - 0.5)));
}
void nb_transfer(const int* rd, int n, int m, std::vector<float>& out, double* rdmad_out = nullptr) {
  double rdmedian = hmedian(rd, (size_t)n);
  const int ns = 31;
  int sub = n / ns;
  std::vector<int> tmp(sub);
  std::vector<double> mads(ns, 0.0);
  for (int j = 0; j < ns; ++j) {
    for (int i = j, k = 0; i < n && k < sub; i += ns, ++k) tmp[k] = (int)std::fabs((float)rd[i] - rdmedian);
    mads[j] = hmedian(tmp.data(), (size_t)sub);
  }
  double rdmad = hmedian(mads.data(), (size_t)ns);
  if (rdmad_out) *rdmad_out = rdmad;
  double r = rdmedian / rdmad;
  int nb = n / m;
  out.assign(nb, 0.f);
  for (int b = 0; b < nb; ++b) {
    int i1 = b * m, i2 = std::min(b * m + m - 1, n - 1);
    double sum = 0.0;
    for (int j = i1; j <= i2; ++j) sum += rd[j];
    out[b] = (float)nb_formula(sum, double(i2 - i1 + 1), r);
  }
  double med_nbt = nb_formula(rdmedian * m, (double)m, r);
  double del_nbt = nb_formula(rdmedian / 2.0 * (double)m, (double)m, r);
  double dup_nbt = nb_formula(rdmedian * 1.5 * (double)m, (double)m, r);
  double tmin = out[0];
  for (int b = 0; b < nb; ++b) if (out[b] < tmin) tmin = out[b];
  for (int b = 0; b < nb; ++b) out[b] = (float)(out[b] - tmin);
  med_nbt -= tmin;
  for (int b = 0; b < nb; ++b) out[b] = (float)(out[b] / med_nbt * rdmedian);
  del_nbt -= tmin; dup_nbt -= tmin;
  del_nbt = del_nbt / med_nbt * rdmedian;
  dup_nbt = dup_nbt / med_nbt * rdmedian;
  med_nbt = med_nbt / med_nbt * rdmedian;
  out[0] = (float)del_nbt; out[1] = (float)dup_nbt; out[2] = (float)med_nbt;  // rsi.cpp:1183-1185
}

// runmeantp(y, float smo, n, band, end_rule=1), wufunctions.cpp:572-647
void runmean(const float* y, float* smo, int n, int band) {
  double sum = 0;
  for (int i = 0; i < band; ++i) sum += (double)y[i];
  double mean = sum / double(band);
  int half = band / 2;
  for (int i = 0; i < half; ++i) smo[i] = (float)mean;
  smo[half] = (float)mean;
  int is = half + 1;
  for (int first = 1, last = band; last < n; ++first, ++last, ++is) {
    sum = sum - (double)y[first - 1] + (double)y[last];
    mean = sum / double(band);
    smo[is] = (float)mean;
  }
  for (int i = is; i < n; ++i) smo[i] = (float)mean;
}

// a15: rsistatus, rsi.cpp:1191-1259 (sequential form, L ascending, first writer wins, 20% break)
void rsistatus(const float* t, const int* medint, int nb, double tmedian, double tlamda, int Lmax, int* status) {
  std::vector<float> rm(nb, 0.f);
  for (int i = 0; i < nb; ++i) status[i] = 0;
  for (int sign = -1; sign <= 1; sign += 2) {  // deletions first, then duplications
    const double lim = sign < 0 ? G.RDmedian * 0.75 : G.RDmedian * 1.25;
    for (int L = 1; L <= Lmax; ++L) {
      std::fill(rm.begin(), rm.end(), 0.f);
      runmean(t, rm.data(), nb, L);
      for (int i = L / 2 + 1; i < nb - L / 2 - 1; ++i) {
        double score = (rm[i] - tmedian) * std::sqrt(double(L));
        if (sign < 0 ? score > -tlamda : score < tlamda) continue;
        int i1 = i - L / 2, i2 = i1 + L - 1;
        double wm = true_median(medint + i1, L);
        if (sign < 0 ? wm > lim : wm < lim) continue;
        if (sign < 0) {
          while (t[i1] > tmedian) i1++;
          while (medint[i1] > lim) i1++;
          while (t[i2] > tmedian) i2--;
          while (medint[i2] > lim) i2--;
        } else {
          while (t[i1] < tmedian) i1++;
          while (medint[i1] < lim) i1++;
          while (t[i2] < tmedian) i2--;
          while (medint[i2] < lim) i2--;
        }
        for (int j = i1; j <= i2; ++j) if (status[j] == 0) status[j] = sign * L;
      }
      int cnt = 0;
      for (int i = 0; i < nb; ++i) cnt += sign < 0 ? (status[i] < 0) : (status[i] > 0);
      if (double(cnt) / double(nb) > 0.2) break;
    }
  }
}

// a17: get_continuous_segments, rsi.cpp:291-326 -- NB the last run is never emitted.
void continuous_segments(const int* status, int nb, int d, std::vector<ocl_cnv>& segs) {
  segs.clear();
  int s0 = 0, s1 = 0, seen = 0;
  for (int i = 0; i < nb; ++i) {
    if (status[i] == 0) continue;
    if (seen == 0) { s0 = s1 = i; seen = 1; continue; }
    if ((double)status[i] * (double)status[s1] > 0 && (i - s1) <= d) { s1 = i; continue; }
    ocl_cnv c = new_cnv(); c.start = s0; c.end = s1;
    segs.push_back(c);
    s0 = s1 = i; ++seen;
  }
}

// a16: filterstatus_tp<float>, rsi.cpp:948-1047 -- per-level means with SEQUENTIAL float sums
void filterstatus(const float* t, int nb, double dev, int* status) {
  int lo = status[0], hi = status[0];
  for (int i = 0; i < nb; ++i) { lo = std::min(lo, status[i]); hi = std::max(hi, status[i]); }
  int nl = hi - lo + 1;
  std::vector<float> sum(nl, 0.0f); std::vector<int> cnt(nl, 0);
  for (int i = 0; i < nb; ++i) { int l = status[i] - lo; sum[l] += t[i]; ++cnt[l]; }
  for (int l = 0; l < nl; ++l) if (cnt[l] != 0) sum[l] /= (double)cnt[l];
  int ldel = lo, ladd = hi;
  for (int l = 0; l < nl; ++l) if (sum[l] < sum[-lo] - dev) { ldel = l + lo; break; }
  for (int l = nl - 1; l >= 0; --l) if (sum[l] > sum[-lo] + dev) { ladd = l + lo; break; }
  if (ldel > 0 || ladd < 0 || ldel > ladd) return;
  double tdel = sum[-lo] - dev, tadd = sum[-lo] + dev;
  std::vector<ocl_cnv> segs;
  continuous_segments(status, nb, 1, segs);
  for (auto& s : segs) {
    int i1 = s.start, i2 = s.end;
    while ((t[i1] > tdel && status[i1] < 0) || (t[i1] < tadd && status[i1] > 0)) { status[i1] = 0; ++i1; if (i1 >= i2) break; }
    while ((t[i2] > tdel && status[i2] < 0) || (t[i2] < tadd && status[i2] > 0)) { status[i2] = 0; --i2; if (i2 <= i1) break; }
  }
}

// a18: get_rsi_segments, rsi.cpp:1060-1117
void rsi_segments(const float* t, const int* status, int nb, double tmedian, std::vector<ocl_cnv>& out) {
  std::vector<ocl_cnv> segs;
  continuous_segments(status, nb, 1, segs);
  for (auto& s : segs) {
    int a = s.start, b = s.end, ns = a, ne = b;
    double best = 0;
    for (int L = 1; L <= b - a + 1; ++L) {
      double sum = 0.0;
      for (int j = a; j <= b && j < a + L; ++j) sum += t[j];
      double sc = std::fabs(sum / (double)L - tmedian) * std::sqrt(double(L));
      if (sc > best) { ns = a; ne = a + L - 1; best = sc; }
      for (int j = a + 1; j + L - 1 <= b; ++j) {
        sum = sum - t[j - 1] + t[j + L - 1];
        sc = std::fabs(sum / (double)L - tmedian) * std::sqrt(double(L));
        if (sc > best) { ns = j; ne = j + L - 1; best = sc; }
      }
    }
    ocl_cnv c = new_cnv();
    c.start = ns; c.end = ne;
    if (hmedian(status + ns, (size_t)(ne - ns + 1)) > 0) { c.type = TYPE_DUP; c.score = best; }
    else { c.type = TYPE_DEL; c.score = -best; }
    out.push_back(c);
  }
}

// a14: rsicnvnbn rsi.cpp:1262-1360 (which=0) / rsicnvmed rsi.cpp:1402-1515 (which=1)
void rsicnv(int which, const float* t, const int* medint, int nb, int* status, std::vector<ocl_cnv>& out) {
  out.clear();
  std::vector<float> tmp(nb);
  double tmedian = which == 0 ? hmedian(t, (size_t)nb) : G.RDmedian;
  for (int i = 0; i < nb; ++i) tmp[i] = (float)std::fabs(t[i] - tmedian);
  double tsigma = hmedian(tmp.data(), (size_t)nb) / 0.6745;
  double tlamda = G.factor * tsigma;
  double target, dev;
  int Lmax = G.Lmax, calmax;
  if (which == 0) {
    target = (t[2] - t[0]) * std::sqrt(2.5);
    tlamda = std::max(tlamda, target);
    double dnb = std::fabs(t[2] - t[0]) + 0.0001;
    calmax = (int)std::pow(tlamda * 2 / dnb, 2);
    dev = tsigma * 3.0;
  } else {
    target = tmedian * std::sqrt(2.0);
    tlamda = std::max(tlamda, target);
    if (G.threshold > 0) tlamda = tmedian * G.threshold;
    calmax = (int)std::pow(tlamda * 4 / (tmedian + 0.001), 2);
    dev = tmedian * 0.6;
  }
  if (Lmax < calmax) Lmax = calmax;
  rsistatus(t, medint, nb, tmedian, tlamda, Lmax, status);
  filterstatus(t, nb, dev, status);
  int k = 0;
  for (int i = 0; i < nb; ++i) if (status[i] == 0) tmp[k++] = t[i];
  if (k > nb / 2) {
    tmedian = hmedian(tmp.data(), (size_t)k);
    for (int i = 0; i < k; ++i) tmp[i] = (float)std::fabs(tmp[i] - tmedian);
    tsigma = hmedian(tmp.data(), (size_t)k) / 0.6745;
    tlamda = std::max(G.factor * tsigma, target);
  }
  if (which == 0) { G.nbnlamda = tlamda; G.nbnmedian = tmedian; } else { G.medlamda = tlamda; G.medmedian = tmedian; }
  rsistatus(t, medint, nb, tmedian, tlamda, Lmax, status);
  std::vector<ocl_cnv> segs;
  rsi_segments(t, status, nb, tmedian, segs);
  for (auto& s : segs) if (!(std::fabs(s.score) < tlamda * 0.5)) out.push_back(s);
}

// a20: isitcnv, rsi.cpp:101-172
void isitcnv(const std::vector<int>& ref, const std::vector<int>& cnv, ocl_cnv& c) {
  int d = (int)cnv.size(), nr = (int)ref.size() - d;
  if (nr <= 0 || d <= 0) { c.status = -9; c.geno = 0; return; }  // the reference aborts here (Array bounds throw)
  std::vector<float> rm(nr);
  double sum = 0;
  for (int i = 0; i < d; ++i) sum += ref[i];
  rm[0] = (float)(sum / double(d));
  for (int i = 1; i < nr; ++i) { sum = sum - ref[i - 1] + ref[i - 1 + d]; rm[i] = (float)(sum / double(d)); }
  double rmed = hmedian(rm.data(), (size_t)nr);
  double rsd = std::sqrt(variance_all(rm.data(), nr));
  if (rsd < 1E-3) rsd = rmed / 40.0 + 1E-3;
  c.length = c.end - c.start + 1;
  c.cnvmed = hmedian(cnv.data(), cnv.size());
  c.cnvsd = std::sqrt(variance_all(cnv.data(), (int)cnv.size()));
  c.cnviqr = hiqr(cnv.data(), cnv.size());
  c.refmed = rmed;
  c.refiqr = hiqr(rm.data(), (size_t)nr);
  c.refsd = c.refiqr / 1.349;
  c.geno = 1; c.status = 1;
  int flag = c.cnvmed > G.RDmedian ? TYPE_DUP : TYPE_DEL;
  if (c.type == TYPE_UNKNOWN) c.type = flag;
  if (c.type != flag) { c.status = -9; return; }
  if (c.type == TYPE_DEL) {
    double reference = std::min(rmed, G.RDmedian);
    reference = std::max(reference, 0.8 * G.RDmedian);
    double nu = (3.0 * c.cnvmed - 2.0 * reference) / rsd;
    c.p1 = pnorm(nu);
    if (nu > 0) { c.status = -9; c.geno = 0; }
  }
  if (c.type == TYPE_DUP) {
    double reference = std::max(rmed, G.RDmedian);
    double nu = (2.5 * c.cnvmed - 3.0 * reference) / rsd / 1.5;
    c.p1 = 1.0 - pnorm(nu);
    if (nu < 0) { c.status = -9; c.geno = 0; }
  }
}

// a20: isitcnvwrap, rsi.cpp:175-287 (neighbour walk spec: SURVEY.md A.1)
void isitcnvwrap(const int* RD, int n, std::vector<ocl_cnv>& list, int ci) {
  const int flag = list[ci].type;
  const int cnvlen = list[ci].end - list[ci].start + 1;
  const int pts = G.maxchkbp * 10;
  const int nl = (int)list.size();
  int d = cnvlen;
  if (n == G.end - G.start + 1) { if (d < G.m * G.minmlen) d = (int)(G.m * G.minmlen); }
  if (n < (G.end - G.start + 1) / 2) { if (d < G.minmlen) d = (int)G.minmlen + 1; }
  std::vector<int> ref((size_t)(int)(G.chklen * d * 2), 0);
  const int buffer = int(cnvlen * G.buffer + 1);
  const double upper = 3.0, lower = 0.15;
  int i = list[ci].start - buffer, idx = ci - 1;
  while (i > 0 && idx > 0 && i < list[idx].start) --idx;
  while (idx > 0 && list[idx].status == -9) --idx;
  int k = (int)(G.chklen * d - 1);
  if (n - list[ci].end < G.chklen * d) k = (int)ref.size() - 1 - n + list[ci].end;
  const int stopper = k;
  while (i > 2 && k >= 0) {
    --i;
    if (flag == TYPE_DEL && RD[i] > G.RDmedian * upper) continue;
    if (flag == TYPE_DUP && RD[i] < G.RDmedian * lower) continue;
    if (idx >= 0 && i >= list[idx].start && i <= list[idx].end) {
      i = list[idx].start - 1; --idx;
      while (idx > 0 && list[idx].status == -9) --idx;
      continue;
    }
    ref[k] = RD[i]; --k;
  }
  if (k >= 0) { int w = 0; for (int k1 = k + 1; k1 <= stopper; ++k1, ++w) ref[w] = ref[k1]; k = w; }
  else k = stopper + 1;
  i = list[ci].end + buffer; idx = ci + 1;
  while (i < n - 2 && idx < nl && i > list[idx].end) ++idx;
  while (idx < nl - 1 && list[idx].status == -9) ++idx;
  while (i < n - 2 && k < 2 * G.chklen * d) {
    ++i;
    if (flag == TYPE_DEL && RD[i] > G.RDmedian * upper) continue;
    if (flag == TYPE_DUP && RD[i] < G.RDmedian * lower) continue;
    if (idx < nl && i >= list[idx].start && i <= list[idx].end) {
      i = list[idx].end + 1; ++idx;
      while (idx < nl - 1 && list[idx].status == -9) ++idx;
      continue;
    }
    ref[k] = RD[i]; ++k;
  }
  if (k < (int)ref.size()) ref.resize(k);
  std::vector<int> cnv(RD + list[ci].start, RD + list[ci].end + 1);
  int tot = (int)ref.size() + (int)cnv.size();
  if (tot > pts) {  // sub-sample, rsi.cpp:264-282
    int dref = (int)((double)ref.size() / (double)tot * (double)pts);
    int dcnv = (int)((double)cnv.size() / (double)tot * (double)pts);
    std::vector<int> r1(dref), c1(dcnv);
    for (int a = 0; a < dref; ++a) r1[a] = ref[(int)(double(a) / double(dref) * double(ref.size()))];
    for (int a = 0; a < dcnv; ++a) c1[a] = cnv[(int)(double(a) / double(dcnv) * double(cnv.size()))];
    ref.swap(r1); cnv.swap(c1);
  }
  isitcnv(ref, cnv, list[ci]);
}

// a19: multisegments rsi.cpp:368-410, areblockscnv rsi.cpp:415-546
void multisegments(const ocl_cnv& seg, const int* status, std::vector<ocl_cnv>& out) {
  out.clear();
  int len = seg.end - seg.start + 1;
  const int* s2 = status + seg.start;
  int lo = s2[0], hi = s2[0];
  for (int i = 0; i < len; ++i) { lo = std::min(lo, s2[i]); hi = std::max(hi, s2[i]); }
  std::vector<int> bin(len);
  for (int lev = lo; lev < hi; ++lev) {
    if (lev == 0) continue;
    int hits = 0;
    for (int i = 0; i < len; ++i) {
      bin[i] = 0;
      if (s2[i] == 0) continue;
      if (s2[i] == lev) ++hits;
      if (lev < 0 && s2[i] < 0 && s2[i] >= lev) bin[i] = 1;
      if (lev > 0 && s2[i] > 0 && s2[i] <= lev) bin[i] = 1;
    }
    if (hits == 0) continue;
    std::vector<ocl_cnv> sub;
    continuous_segments(bin.data(), len, 1, sub);
    for (auto& s : sub) {
      s.start = std::max(s.start + seg.start, seg.start);
      s.end = std::min(s.end + seg.start, seg.end);
      out.push_back(s);
    }
  }
}
void areblockscnv(const int* medint, const int* status, int nb, std::vector<ocl_cnv>& list) {
  for (int i = 0; i < (int)list.size(); ++i) isitcnvwrap(medint, nb, list, i);
  for (int i = 0; i < (int)list.size(); ++i) {
    if (list[i].status != -9) continue;
    if (list[i].type == TYPE_DEL && list[i].cnvmed < 0.7 * list[i].refmed) { list[i].geno = 1; list[i].p1 = G.p; continue; }
    if (list[i].type == TYPE_DUP && list[i].cnvmed > 1.3 * list[i].refmed) { list[i].geno = 1; list[i].p1 = G.p; continue; }
    ocl_cnv orig = list[i], pick = list[i];
    std::vector<ocl_cnv> sub;
    multisegments(pick, status, sub);
    for (int j = (int)sub.size() - 1; j >= 0; --j) {
      sub[j].type = pick.type;
      list[i] = sub[j];
      isitcnvwrap(medint, nb, list, i);
      sub[j] = list[i];
    }
    for (int j = (int)sub.size() - 1; j >= 0; --j) {
      if (sub[j].geno == 0) continue;
      if (pick.geno == 0) pick = sub[j];
      if (sub[j].length > pick.length) pick = sub[j];
    }
    if (pick.geno == 0) pick = orig;
    list[i] = pick;
  }
}

// a22: sortcnvstartposition, rsi.cpp:549-577 (stable by start)
void sort_by_start(std::vector<ocl_cnv>& list) {
  for (auto& c : list) if (c.start > c.end) std::swap(c.start, c.end);
  std::stable_sort(list.begin(), list.end(), [](const ocl_cnv& a, const ocl_cnv& b) { return a.start < b.start; });
}

// a21: optimize_with_derivative, rsi.cpp:889-944
void optimize_one(const int* RD, int n, ocl_cnv& c) {
  int len = c.end - c.start + 1;
  int disp = std::max(250, len / 4);
  int ns = c.start - disp, ne = c.end + disp;
  if (ns < 2 * len) return;
  if (ne > n - 2 * len) return;
  std::vector<double> dd;
  double diff = 0.0;
  for (int k = ns - len; k < ns; ++k) diff += RD[k];
  for (int k = ns; k < ns + len; ++k) diff -= RD[k];
  dd.push_back(diff);
  for (int i = ns + 1; i < ne; ++i) { diff = diff - RD[i - 1 - len] + RD[i - 1] + RD[i - 1] - RD[i - 1 + len]; dd.push_back(diff); }
  int imax = -1; double best = 0;
  for (int i = 0; i < 2 * disp; ++i) {
    if (c.type == TYPE_DEL && dd[i] > best) { best = dd[i]; imax = i; }
    if (c.type == TYPE_DUP && dd[i] < best) { best = dd[i]; imax = i; }
  }
  if (imax > 0) c.start = ns + imax;
  imax = -1; best = 0;
  for (int i = (int)dd.size() - 2 * disp; i < (int)dd.size(); ++i) {
    if (c.type == TYPE_DEL && dd[i] < best) { best = dd[i]; imax = i; }
    if (c.type == TYPE_DUP && dd[i] > best) { best = dd[i]; imax = i; }
  }
  if (imax > 0) c.end = ne - (int)dd.size() + imax;
}

// a23: mergesegments, rsi.cpp:694-885
void mergesegments(const int* RD, int n, std::vector<ocl_cnv>& list) {
  std::vector<ocl_cnv> test;
  for (int i = 0; i < (int)list.size() - 1; ++i) {
    if (list[i].type != list[i + 1].type) continue;
    if (!(std::max(list[i].start, list[i + 1].start) < std::min(list[i].end, list[i + 1].end))) continue;
    ocl_cnv u = list[i];
    u.start = std::min(list[i].start, list[i + 1].start);
    u.end = std::max(list[i].end, list[i + 1].end);
    test = list; test[i] = u; test[i + 1] = u; test[i + 1].status = -9;
    isitcnvwrap(RD, n, test, i);
    if (test[i].geno == 0) {
      test = list; test[i + 1].status = -9;
      isitcnvwrap(RD, n, test, i);
      test[i].status = -9; test[i + 1].status = 0;
      isitcnvwrap(RD, n, test, i + 1);
      if (test[i + 1].p1 < test[i].p1) test[i] = test[i + 1];
      if (test[i].p1 > G.p) { list[i].status = -9; list[i + 1].status = -9; }
    }
    if (test[i].geno == 0) continue;
    list[i] = test[i]; list[i].status = -9;
    list[i + 1] = test[i]; list[i + 1].status = 0;
  }
  std::vector<ocl_cnv> keep;
  for (auto& c : list) if (c.status != -9) keep.push_back(c);
  list = keep;
  if (!G.merge) return;
  for (int i = 0; i < (int)list.size() - 1; ++i) {
    if (list[i].type != list[i + 1].type) continue;
    if (list[i].geno == 0 || list[i + 1].geno == 0) continue;
    int gap = list[i + 1].start - list[i].end;
    if (gap > (list[i].end - list[i].start) * G.chklen * 0.7 && gap > (list[i + 1].end - list[i + 1].start) * G.chklen * 0.7) continue;
    double m1 = mean_range(RD, list[i].start, list[i].end);
    double m2 = mean_range(RD, list[i + 1].start, list[i + 1].end);
    double cm = (m1 * (list[i].end - list[i].start) + m2 * (list[i + 1].end - list[i + 1].start)) /
                ((list[i].end - list[i].start) + (list[i + 1].end - list[i + 1].start));
    double mm = mean_range(RD, list[i].start, list[i + 1].end);
    if (list[i].type == TYPE_DEL && mm > cm + 1.5 * list[i + 1].refsd + 1.5 * list[i].refsd) continue;
    if (list[i].type == TYPE_DUP && mm < cm - 1.5 * list[i + 1].refsd - 1.5 * list[i].refsd) continue;
    ocl_cnv u = list[i];
    u.start = list[i].start; u.end = list[i + 1].end;
    test = list; test[i] = u; test[i + 1] = u; test[i + 1].status = -9;
    isitcnvwrap(RD, n, test, i);
    if (test[i].geno == 0) continue;
    list[i] = test[i]; list[i + 1] = test[i]; list[i].status = -9;
  }
  keep.clear();
  for (auto& c : list) if (c.status != -9) keep.push_back(c);
  list = keep;
}

// expand_coordinate, rsi.cpp:1524-1551
int expand_coordinate(int p) {
  if (G.nbeg.empty()) return p;
  int dx = 0;
  std::vector<int> brk, inc;
  for (size_t i = 0; i < G.nbeg.size(); ++i) { dx += G.nend[i] - G.nbeg[i] + 1; brk.push_back(G.nend[i] + 1 - dx); inc.push_back(dx); }
  if (p < brk[0]) return p;
  if (p >= brk.back()) return p + inc.back();
  for (size_t i = 0; i + 1 < brk.size(); ++i) if (p >= brk[i] && p < brk[i + 1]) return p + inc[i];
  return p;
}

// a25: sd_filters, rsi.cpp:1753-1792
void sd_filters(std::vector<ocl_cnv>& list) {
  int minlen = std::max(G.m * 2, 500);
  double tsd = G.RDsd / 1.2;
  std::vector<ocl_cnv> keep;
  for (auto& c : list) {
    bool k = true;
    int span = std::abs(c.end - c.start);
    if (span < 1000) k = false;
    if (c.type == TYPE_DEL) {
      if (c.p1 > 0.2) k = false;
      if (c.refsd > 0.6 * tsd) k = false;
      if (c.cnvsd > 1.3 * tsd) k = false;
      if (c.cnvsd * G.RDmedian > 2.5 * c.cnvmed * tsd) k = false;
      if (c.cnvmed < 0.66 * std::min(G.RDmedian, c.refmed) && c.cnvsd < tsd && span > 800) k = true;
    }
    if (c.type == TYPE_DUP) {
      if (c.p1 > 0.05) k = false;
      if (c.refsd > 0.6 * tsd) k = false;
      if (c.cnvsd * G.RDmedian > 2.0 * c.cnvmed * tsd) k = false;
    }
    if (span < minlen) k = false;
    if (k) keep.push_back(c);
  }
  list = keep;
}

// L3: detectcnv, rsi.cpp:1795-1945.  Optionally exports the bin-level intermediates.
struct BinDump { std::vector<float> med, nbn; std::vector<int> medint, status; };
// the intermediate call lists of the last detectcnv (parity of RSIGPU_ARR_SEGMENTS/BLOCKS/PREMERGE/MERGED):
// [0] segments leaving rsicnvnbn/rsicnvmed (the last transformation run), [1] after areblockscnv, [2] after
// bins->bases + 2x optimize + sort, [3] after mergesegments + sort
std::vector<ocl_cnv> g_lists[4];
void detectcnv(const int* RD, int n, std::vector<ocl_cnv>& out, BinDump* dump) {
  out.clear();
  for (auto& l : g_lists) l.clear();
  if (G.RDmedian < 5) return;
  const int m = G.m;
  std::vector<float> med, nbn;
  median_transfer(RD, n, m, med);
  int nb = (int)med.size();
  std::vector<int> medint(nb);
  for (int i = 0; i < nb; ++i) medint[i] = (int)(med[i] + 0.5);
  G.RDmedian = hmedian(RD, (size_t)n);
  nb_transfer(RD, n, m, nbn);
  G.factor = std::sqrt(2.0 * (1.0 + G.epsilon) * std::log(3.1E9));
  G.Lmax = std::max(10000 / m, 20);
  std::vector<int> st_med(nb, 0), st_nbn(nb, 0);
  std::vector<ocl_cnv> segs;
  if (G.trans != 0) { rsicnv(1, med.data(), medint.data(), nb, st_med.data(), segs); g_lists[0] = segs; areblockscnv(medint.data(), st_med.data(), nb, segs); }
  if (G.trans == 0) { rsicnv(0, nbn.data(), medint.data(), nb, st_nbn.data(), segs); g_lists[0] = segs; areblockscnv(medint.data(), st_nbn.data(), nb, segs); }
  if (G.trans == 2) {
    std::vector<ocl_cnv> s2;
    rsicnv(0, nbn.data(), medint.data(), nb, st_nbn.data(), s2); g_lists[0] = s2; areblockscnv(medint.data(), st_nbn.data(), nb, s2);
    segs.insert(segs.end(), s2.begin(), s2.end());
  }
  g_lists[1] = segs;
  if (dump) { dump->med = med; dump->nbn = nbn; dump->medint = medint; dump->status = G.trans == 1 ? st_med : st_nbn; }
  sort_by_start(segs);
  std::vector<ocl_cnv> list;
  for (auto& s : segs) {
    if (s.geno == 0) continue;
    if (s.start == s.end) continue;
    s.start = s.start * m + m / 2;
    s.end = s.end * m + m / 2;
    if (s.start < 0) s.start = 0;
    if (s.end > n - 1) s.end = n - 1;
    s.length = s.end - s.start + 1;
    list.push_back(s);
  }
  for (auto& c : list) c.tid = G.tid;
  for (int rep = 0; rep < 2; ++rep) for (auto& c : list) optimize_one(RD, n, c);
  sort_by_start(list);
  g_lists[2] = list;
  mergesegments(RD, n, list);
  sort_by_start(list);
  g_lists[3] = list;
  for (int i = 0; i < (int)list.size(); ++i) {
    double len = double(list[i].end - list[i].start + 1) / double(m);
    isitcnvwrap(RD, n, list, i);
    list[i].score = (list[i].cnvmed - G.RDmedian) * std::sqrt(len);
    int p1 = expand_coordinate(list[i].start), p2 = expand_coordinate(list[i].end);
    for (size_t k = 0; k < G.nbeg.size(); ++k) if (std::max(p1, G.nbeg[k]) <= std::min(p2, G.nend[k])) list[i].status = -9;
    if (list[i].status != -9) out.push_back(list[i]);
  }
  for (auto& c : out) { c.start = expand_coordinate(c.start); c.end = expand_coordinate(c.end); }
}

// ---------------------------------------------------------------------------------------------
// BAM path.  Reads arrive as the same SoA the C-ABI takes (include/rsigpu.h).
struct Reads {
  int64_t n; int tid;
  const int32_t *pos, *mpos, *isize, *mtid;
  const uint16_t* flag; const uint8_t* mapq;
  const uint32_t* cigar_off; const uint32_t* cigar;
  const uint64_t* qual_off; const uint8_t* qual;
};
enum { F_PROPER = 2, F_REV = 16, F_MREV = 32, F_SECONDARY = 256, F_DUP = 1024 };

uint32_t calend(const Reads& R, int64_t r) {  // bam_calend, samtools-0.1.18/bam.c:17-27
  uint32_t end = (uint32_t)R.pos[r];
  for (uint32_t k = R.cigar_off[r]; k < R.cigar_off[r + 1]; ++k) {
    int op = R.cigar[k] & 15;
    if (op == 0 || op == 2 || op == 3) end += R.cigar[k] >> 4;
  }
  return end;
}

// a5: load_data_from_bam hot loop loaddata.cpp:312-335 + resolve_cigar_pos samfunctions.cpp:38-100
void pileup(const Reads& R, int* RD, int L) {
  for (int64_t r = 0; r < R.n; ++r) {
    if (R.pos[r] == 0) continue;
    if (R.mapq[r] < G.minq) continue;
    if (R.flag[r] & F_SECONDARY) continue;
    if (R.flag[r] & F_DUP) continue;
    uint32_t c0 = R.cigar_off[r], c1 = R.cigar_off[r + 1];
    int nc = (int)(c1 - c0);
    int anchor = -1;
    std::vector<uint32_t> qop(nc), cop(nc, 0);
    uint32_t q = 0;
    for (int k = 0; k < nc; ++k) {
      int op = R.cigar[c0 + k] & 15; uint32_t l = R.cigar[c0 + k] >> 4;
      qop[k] = q;
      if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) q += l;
      if ((op == 0 || op == 2 || op == 7 || op == 8) && anchor < 0) anchor = k;
    }
    if (anchor < 0) continue;  // no M/D/=/X: cop stays 0 and nothing is an M or = op
    uint32_t e = (uint32_t)R.pos[r] + 1;
    for (int k = anchor; k < nc; ++k) {
      int op = R.cigar[c0 + k] & 15;
      cop[k] = e;
      if (op == 0 || op == 2 || op == 3 || op == 4) e += R.cigar[c0 + k] >> 4;
    }
    e = (uint32_t)R.pos[r] + 1;
    for (int k = anchor - 1; k >= 0; --k) {
      int op = R.cigar[c0 + k] & 15;
      if (op == 0 || op == 2 || op == 3 || op == 4) e -= R.cigar[c0 + k] >> 4;
      cop[k] = e;
    }
    const uint8_t* qual = R.qual + R.qual_off[r];
    for (int k = 0; k < nc; ++k) {
      int op = R.cigar[c0 + k] & 15; uint32_t l = R.cigar[c0 + k] >> 4;
      if (op != 0 && op != 7) continue;
      int p1 = (int)cop[k] - 1; int q1 = (int)qop[k];
      for (uint32_t i = 0; i < l && p1 < L; ++i, ++p1, ++q1) if (qual[q1] >= G.min_baseQ) ++RD[p1];
    }
  }
}

// a26: insert-size sample of bam_rd_pr_stats, pairrd.cpp:112-260 (order of tests: SURVEY.md A.2)
void isize_stats(const Reads& R, int tid_len, int* isize_mean, int* isize_sd) {
  const uint32_t beg = 10000000u, end = 349250621u;
  double s = 0, s2 = 0, c = 0;
  size_t count = 0; int pos_start = 0, pos_end = -10000;
  for (int64_t r = 0; r < R.n; ++r) {
    if ((uint32_t)R.pos[r] >= end) break;                       // bam_iter_read stop rule, bam_index.c:691
    uint32_t re = R.cigar_off[r + 1] > R.cigar_off[r] ? calend(R, r) : (uint32_t)R.pos[r] + 1;
    if (!(re > beg && (uint32_t)R.pos[r] < end)) continue;      // is_overlap, bam_index.c:564-569
    if (R.mtid[r] != R.tid && R.mtid[r] > 0) continue;
    if (R.flag[r] & F_SECONDARY) continue;
    if (R.flag[r] & F_DUP) continue;
    if ((R.flag[r] & F_PROPER) && R.mtid[r] == R.tid) {
      s += std::abs(R.isize[r]);
      s2 += (int32_t)((uint32_t)R.isize[r] * (uint32_t)R.isize[r]);
      c += 1;
    }
    int rpe = (int)calend(R, r);
    if (R.pos[r] >= tid_len) break;
    if (rpe >= tid_len) break;
    if (R.pos[r] > pos_end + 1000) { count = 0; pos_start = R.pos[r]; pos_end = R.pos[r]; }
    pos_end = R.pos[r];
    count++;
    if (count > 1000000 || (pos_end - pos_start) > 1000000) break;
  }
  *isize_mean = -1; *isize_sd = -1;
  if (c > 2) {
    s /= c;
    double sd = std::sqrt((s2 - c * s * s) / c);
    *isize_mean = (int)s; *isize_sd = (int)sd;
  }
}

// a27: cnv_stat, pairrd.cpp:622-748
void cnv_stat(const Reads& R, int tid_len, std::vector<ocl_cnv>& list) {
  if (list.empty()) return;
  int im, isd;
  isize_stats(R, tid_len, &im, &isd);
  int DIS = 1000;
  for (auto& c : list) {
    int beg = c.start, end = c.end;
    if (beg > end) std::swap(beg, end);
    int LEN = end - beg + 1;
    DIS = std::max(DIS, LEN); DIS = std::min(DIS, 5000);
    int p1e = beg - DIS, p2e = end + DIS;
    if (p1e < 1) p1e = 1;
    double qall = 0, q0 = 0; size_t rp = 0;
    for (int64_t r = 0; r < R.n; ++r) {
      if ((uint32_t)R.pos[r] >= (uint32_t)p2e) break;
      int nc = (int)(R.cigar_off[r + 1] - R.cigar_off[r]);
      uint32_t re = nc ? calend(R, r) : (uint32_t)R.pos[r] + 1;
      if (!(re > (uint32_t)p1e && (uint32_t)R.pos[r] < (uint32_t)p2e)) continue;
      if (nc <= 1) continue;
      int rbeg = R.pos[r], rend = (int)re;
      if (rend > beg && rbeg < end) { qall += 1; if (R.mapq[r] == 0) q0 += 1; }
      if (R.mtid[r] != R.tid && R.mtid[r] > 0) continue;
      int F = R.flag[r];
      if ((F & F_REV) == 0 && (F & F_MREV) == 0) continue;
      if ((F & F_REV) > 0 && (F & F_MREV) > 0) continue;
      int r1 = rend, r2 = R.mpos[r];
      if (c.type == TYPE_DEL) {
        if (r2 - r1 < im + isd * 3) continue;
        int ov = std::min(r2, end) - std::max(r1, beg);
        if (ov < 0) continue;
        if (std::abs(r1 - beg) + std::abs(r2 - end) < im + isd * 3) { ++rp; continue; }
        if (ov < LEN * 0.5) continue;
        if (ov < (r2 - r1) * 0.5) continue;
        ++rp; continue;
      }
      if (c.type == TYPE_DUP) {
        if (r2 - r1 > im - isd * 3) continue;
        if (std::abs(r1 - beg) + std::abs(r2 - end) < im + isd * 3) { ++rp; continue; }
        if (r1 > r2) std::swap(r1, r2);
        int ov = std::min(r2, end) - std::max(r1, beg);
        if (ov < LEN * 0.5) continue;
        if (ov < (r2 - r1) * 0.5) continue;
        ++rp; continue;
      }
    }
    c.q0 = q0 / (qall + 0.00001);
    c.rp = (int)rp;
  }
}

int list_out(const std::vector<ocl_cnv>& v, ocl_cnv* out, int cap) {
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

void ocl_set_params(int m, int minq, int min_baseQ, double cap, int gcadjust, int trans, int merge,
                    double threshold, double epsilon) {
  G = Ctx();
  if (m % 2 != 1) m += 1;  // rsi.cpp:2061-2064
  G.m = m; G.minq = minq; G.min_baseQ = min_baseQ; G.cap = cap; G.gcadjust = gcadjust != 0; G.trans = trans;
  G.merge = merge != 0; G.threshold = threshold; G.epsilon = epsilon;
}
// the undocumented knobs -reflen / -maxchkbp (rsi.cpp:2024-2026); call after ocl_set_params
void ocl_set_knobs(double chklen, int maxchkbp) { G.chklen = chklen; G.maxchkbp = maxchkbp; }
void ocl_set_state(double RDmedian, double RDsd, int start, int end, int Lmax, double factor) {
  G.RDmedian = RDmedian; G.RDsd = RDsd; G.start = start; G.end = end; G.Lmax = Lmax; G.factor = factor;
}
void ocl_get_state(double* o) {
  o[0] = G.RDmedian; o[1] = G.RDsd; o[2] = G.start; o[3] = G.end; o[4] = G.Lmax; o[5] = G.factor;
  o[6] = G.nbnmedian; o[7] = G.nbnlamda; o[8] = G.medmedian; o[9] = G.medlamda;
}
void ocl_set_noncode(const int* b, const int* e, int n) { G.nbeg.assign(b, b + n); G.nend.assign(e, e + n); }
int ocl_get_noncode(int* b, int* e, int cap) {
  for (int i = 0; i < (int)G.nbeg.size() && i < cap; ++i) { b[i] = G.nbeg[i]; e[i] = G.nend[i]; }
  return (int)G.nbeg.size();
}

double ocl_median_i32(const int* x, long n) { return hmedian(x, (size_t)n); }
double ocl_median_f32(const float* x, long n) { return hmedian(x, (size_t)n); }
double ocl_median_f64(const double* x, long n) { return hmedian(x, (size_t)n); }
double ocl_iqr_i32(const int* x, long n) { return hiqr(x, (size_t)n); }
double ocl_iqr_f32(const float* x, long n) { return hiqr(x, (size_t)n); }
double ocl_true_median_i32(const int* x, long n) { return true_median(x, (size_t)n); }
double ocl_pnorm(double x) { return pnorm(x); }
double ocl_variance_i32(const int* x, int n) { return variance_all(x, n); }
double ocl_variance_f32(const float* x, int n) { return variance_all(x, n); }
void ocl_runmean_f32(const float* y, float* smo, int n, int band) { runmean(y, smo, n, band); }

int ocl_noseq_regions(const uint8_t* fasta, int n, int* b, int* e, int cap) {
  noseq_regions(fasta, n, G.m, G.nbeg, G.nend);
  return ocl_get_noncode(b, e, cap);
}
void ocl_checkgccontent(int* rd, const uint8_t* gc, int n) { gc_adjust(rd, gc, n); }
double ocl_apply_cap(int* rd, int n) { if (G.cap <= 1) return G.RDmedian; G.RDmedian = apply_cap(rd, n, G.cap); return G.RDmedian; }
int ocl_concatenate(int* rd, int n) {
  int w = compact(rd, n, G.nbeg, G.nend);
  if (!G.nbeg.empty()) { G.start = 1; G.end = w; }
  return w;
}
int ocl_median_transfer(const int* rd, int n, int m, float* out) {
  std::vector<float> v; median_transfer(rd, n, m, v); std::copy(v.begin(), v.end(), out); return (int)v.size();
}
int ocl_nb_transfer(const int* rd, int n, int m, float* out) {
  std::vector<float> v; nb_transfer(rd, n, m, v); std::copy(v.begin(), v.end(), out); return (int)v.size();
}
void ocl_rsistatus(const float* t, const int* medint, int nb, double tmedian, double tlamda, int Lmax, int* status) {
  rsistatus(t, medint, nb, tmedian, tlamda, Lmax, status);
}
void ocl_filterstatus(const float* t, int nb, double dev, int* status) { filterstatus(t, nb, dev, status); }
int ocl_continuous_segments(const int* status, int nb, int d, ocl_cnv* out, int cap) {
  std::vector<ocl_cnv> v; continuous_segments(status, nb, d, v); return list_out(v, out, cap);
}
int ocl_get_rsi_segments(const float* t, const int* status, int nb, double tmedian, ocl_cnv* out, int cap) {
  std::vector<ocl_cnv> v; rsi_segments(t, status, nb, tmedian, v); return list_out(v, out, cap);
}
int ocl_rsicnv(int which, const float* t, const int* medint, int nb, int* status, ocl_cnv* out, int cap) {
  std::vector<ocl_cnv> v; rsicnv(which, t, medint, nb, status, v); return list_out(v, out, cap);
}
void ocl_isitcnvwrap(const int* rd, int n, ocl_cnv* list, int nlist, int idx) {
  std::vector<ocl_cnv> v(list, list + nlist); isitcnvwrap(rd, n, v, idx); list_out(v, list, nlist);
}
int ocl_areblockscnv(const int* medint, const int* status, int nb, ocl_cnv* list, int nlist) {
  std::vector<ocl_cnv> v(list, list + nlist); areblockscnv(medint, status, nb, v); return list_out(v, list, nlist);
}
void ocl_sort(ocl_cnv* list, int nlist) { std::vector<ocl_cnv> v(list, list + nlist); sort_by_start(v); list_out(v, list, nlist); }
void ocl_optimize(const int* rd, int n, ocl_cnv* list, int nlist) { for (int i = 0; i < nlist; ++i) optimize_one(rd, n, list[i]); }
int ocl_mergesegments(const int* rd, int n, ocl_cnv* list, int nlist) {
  std::vector<ocl_cnv> v(list, list + nlist); mergesegments(rd, n, v); return list_out(v, list, nlist);
}
int ocl_sd_filters(ocl_cnv* list, int nlist) { std::vector<ocl_cnv> v(list, list + nlist); sd_filters(v); return list_out(v, list, nlist); }
int ocl_expand_coordinate(int p) { return expand_coordinate(p); }
int ocl_last_list(int which, ocl_cnv* out, int cap) {
  if (which < 0 || which > 3) return -1;
  return list_out(g_lists[which], out, cap);
}
int ocl_detectcnv(const int* rd, int n, ocl_cnv* out, int cap) {
  std::vector<ocl_cnv> v; detectcnv(rd, n, v, nullptr); return list_out(v, out, cap);
}

// Whole depth path after text parsing (loaddata.cpp:481-538, rsi.cpp:2200-2208); same contract as
// ref_depth_path in oracle/ref_harness.cpp.  Optional bin-level dumps (each nb long, may be NULL).
int ocl_depth_path(int* depth, const uint8_t* fasta, int n, int stage, int* n_compact, double* chr_stats,
                   ocl_cnv* out, int cap, float* bin_med, float* bin_nbn, int* bin_medint, int* bin_status) {
  std::vector<uint8_t> gc(n);
  for (int k = 0; k < n; ++k) gc[k] = (fasta[k] == 'G' || fasta[k] == 'C');
  noseq_regions(fasta, n, G.m, G.nbeg, G.nend);
  G.start = 1; G.end = n;
  if (G.gcadjust) gc_adjust(depth, gc.data(), n);
  if (G.cap > 1) G.RDmedian = apply_cap(depth, n, G.cap);
  *n_compact = n;
  int w = n;
  if (stage >= 1) {
    w = compact(depth, n, G.nbeg, G.nend);
    if (!G.nbeg.empty()) { G.start = 1; G.end = w; }
    G.RDmedian = hmedian(depth, (size_t)w);
    G.RDsd = std::sqrt(variance_all(depth, w));
    *n_compact = w;
  }
  std::vector<ocl_cnv> v;
  BinDump dump;
  if (stage >= 2) detectcnv(depth, w, v, &dump);
  if (stage >= 3) sd_filters(v);
  chr_stats[0] = G.RDmedian; chr_stats[1] = G.RDsd;
  if (stage >= 2) {
    if (bin_med) std::copy(dump.med.begin(), dump.med.end(), bin_med);
    if (bin_nbn) std::copy(dump.nbn.begin(), dump.nbn.end(), bin_nbn);
    if (bin_medint) std::copy(dump.medint.begin(), dump.medint.end(), bin_medint);
    if (bin_status) std::copy(dump.status.begin(), dump.status.end(), bin_status);
  }
  return list_out(v, out, cap);
}

// BAM path pieces on the read SoA
void ocl_pileup(int64_t n, int tid, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq,
                const uint32_t* cigar_off, const uint32_t* cigar, const uint64_t* qual_off, const uint8_t* qual,
                int* rd, int L) {
  Reads R{n, tid, pos, nullptr, nullptr, nullptr, flag, mapq, cigar_off, cigar, qual_off, qual};
  pileup(R, rd, L);
}
void ocl_isize_stats(int64_t n, int tid, const int32_t* pos, const int32_t* mpos, const int32_t* isize, const int32_t* mtid,
                     const uint16_t* flag, const uint32_t* cigar_off, const uint32_t* cigar, int tid_len, int* out2) {
  Reads R{n, tid, pos, mpos, isize, mtid, flag, nullptr, cigar_off, cigar, nullptr, nullptr};
  isize_stats(R, tid_len, &out2[0], &out2[1]);
}
void ocl_cnv_stat(int64_t n, int tid, const int32_t* pos, const int32_t* mpos, const int32_t* isize, const int32_t* mtid,
                  const uint16_t* flag, const uint8_t* mapq, const uint32_t* cigar_off, const uint32_t* cigar,
                  int tid_len, ocl_cnv* list, int nlist) {
  Reads R{n, tid, pos, mpos, isize, mtid, flag, mapq, cigar_off, cigar, nullptr, nullptr};
  std::vector<ocl_cnv> v(list, list + nlist);
  cnv_stat(R, tid_len, v);
  list_out(v, list, nlist);
}

// row formatting, cnv_format1 rsi.cpp:581-631 (default ostream formatting = %g with 6 digits)
int ocl_format_row(const ocl_cnv* c, const char* chrom, double rdmedian, double rdsd, char* buf, int cap) {
  static const char* T[] = {"DEL", "DUP", "UNKNOWN"};
  double q1 = c->p1 < 1.0E-10 ? 99 : -10.0 * std::log(c->p1) / std::log(10.0);
  return std::snprintf(buf, cap, "%s\t%d\t%d\t%s\t%d\t%d\t%g(%g);%g(%g);%g(%g)\tRP=%d;Q0=%g\trsi", chrom, c->start, c->end,
                       T[c->type], (int)q1, c->end - c->start + 1, c->cnvmed, c->cnviqr / 1.349, c->refmed,
                       c->refiqr / 1.349, rdmedian, rdsd, c->rp, c->q0);
}

}  // extern "C"
