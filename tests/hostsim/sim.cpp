// TEST-ONLY build of rsicnv_b200/csrc/candidates.cuh with a plain C++ compiler: `Cta` degenerates to a
// one-thread block (see cta.cuh), so the candidate-stage control flow can be checked against the
// oracle on a machine without a GPU.  Never linked into the product.
#include <cstdlib>
#include <vector>
#include "../../rsicnv_b200/csrc/candidates.cuh"
using namespace rsigpu;

namespace {
struct Sim {
  CandCfg P; CandScratch S; std::vector<int> ref, sub; std::vector<long long> pref; std::vector<float> rm; std::vector<unsigned> hist; int err = 0;
  double bc[32];
  Cta cta;
  Sim(int n) {
    int cap = n + 16;
    ref.resize(cap); sub.resize(2000016); pref.resize(cap + 1); rm.resize(cap); hist.resize(1 << 22);
    S.ref = ref.data(); S.ref_cap = cap; S.sub = sub.data(); S.sub_cap = (int)sub.size(); S.pref = pref.data(); S.rm = rm.data();
    S.hist = hist.data(); S.hist_cap = (int)hist.size(); S.shist = nullptr; S.shist_cap = 0; S.err = &err; S.prof = nullptr;
    cta.bc = bc;
  }
};
CandCfg g_cfg;
}

extern "C" {
void sim_set_cfg(int m, int maxchkbp, int merge, int tid, double chklen, double rdmedian, double rdsd, int span) {
  g_cfg.m = m; g_cfg.maxchkbp = maxchkbp; g_cfg.merge = merge; g_cfg.tid = tid; g_cfg.chklen = chklen; g_cfg.minmlen = 3.01;
  g_cfg.buffer = 0.05; g_cfg.p = 0.05; g_cfg.rdmedian = rdmedian; g_cfg.rdsd = rdsd; g_cfg.span = span;
}
double sim_phi(double x) { return phi(x); }
int sim_isitcnvwrap(const int* rd, int n, Cnv* list, int nl, int idx) {
  Sim s(n); s.P = g_cfg;
  cnv_test(s.cta, s.P, s.S, rd, n, plain_view(list, nl), idx, &list[idx]);
  return s.err;
}
int sim_areblockscnv(const int* medint, const int* status, int nb, Cnv* list, int nl) {
  Sim s(nb); s.P = g_cfg; Cnv ov[2];
  blocks_test(s.cta, s.P, s.S, medint, status, nb, list, nl, ov);
  return nl;
}
void sim_sort(Cnv* list, int nl) { Sim s(16); std::vector<Cnv> tmp(nl + 1); list_sort(s.cta, list, nl, tmp.data()); }
void sim_optimize(const int* rd, int n, Cnv* list, int nl) { Sim s(16); for (int j = 0; j < nl; ++j) edge_refine(s.cta, rd, n, &list[j]); }
int sim_mergesegments(const int* rd, int n, Cnv* list, int nl) {
  Sim s(n); s.P = g_cfg; Cnv ov[2];
  return merge_segments(s.cta, s.P, s.S, rd, n, list, nl, ov);
}
int sim_sd_filters(Cnv* list, int nl) { return sd_filter_list(g_cfg, list, nl); }
int sim_expand(int p, const int* nbeg, const int* nend, int nn) { return expand_coord(p, nbeg, nend, nn); }
double sim_median_f32(const float* x, int n) { Sim s(16); double q[3]; cta_hist_stat(s.cta, s.S, x, n, 0.01, q); return q[1]; }
double sim_median_i32(const int* x, int n) { Sim s(16); double q[3]; cta_hist_stat(s.cta, s.S, x, n, 1.0, q); return q[1]; }
double sim_iqr_f32(const float* x, int n) { Sim s(16); double q[3]; cta_hist_stat(s.cta, s.S, x, n, 0.01, q); return q[2] - q[0]; }
// detectcnv from areblockscnv onwards; list holds rsi segments in bin coordinates; returns final count
int sim_candidates(const int* rd, int n, const int* medint, const int* status, int nb, const int* nbeg, const int* nend, int nn,
                   Cnv* list, int nl, int cap, int* err) {
  Sim s(n); s.P = g_cfg; Cnv ov[2]; std::vector<Cnv> tmp(cap + 1);
  CandDumps D{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  int r = candidates_main(s.cta, s.P, s.S, rd, n, medint, status, nb, nbeg, nend, nn, list, nl, tmp.data(), ov, D, 0);
  *err = s.err;
  return r;
}
}
