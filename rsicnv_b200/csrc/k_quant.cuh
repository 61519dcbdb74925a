// k_quant.cuh -- histogram ("partition") quantiles of a float bin array, the way the reference's
// _median / _interquartilerange macros compute them (partition_stat_tp, wufunctions.cpp:363-424,
// dy = 0.01 for float): bucket (size_t)((x - ymin)/dy + 0.5), value ymin + b*dy at the first bucket
// whose running count reaches n/4, n/2, 3n/4; the MEAN when the range is below dy.
//
// Three launches per evaluation: k_fq_minmax -> k_fq_hist -> k_fq_pick.  A job slot of DevState is
// zero-initialised once per contig (omin is stored complemented so that zero works); the global
// bucket array is left zeroed again by k_fq_pick.  The scalar computations that follow a quantile in
// rsicnvnbn / rsicnvmed (rsi.cpp:1271-1290, 1410-1430, 1312-1318) are one-thread kernels in k_rsi.cuh,
// so no host round trip is needed.
#pragma once
#include "k_bins.cuh"

namespace rsigpu {

enum { QM_ID = 0, QM_ABSDEV = 1 };
enum { FQ_SMEM_BINS = 8192 };

__device__ __forceinline__ float fq_value(const float* __restrict__ x, int i, int mode, double center) {
  const float v = x[i];
  if (mode == QM_ABSDEV) return (float)fabs((double)v - center);
  return v;
}

__global__ void k_fq_minmax(const float* __restrict__ x, const int* __restrict__ status, int masked, int mode, const double* center_p,
                            DevState* st, int slot) {
  RSI_CTA_SETUP(c);
  const int n = st->nb;
  const double center = center_p ? *center_p : 0.0;
  u32 mn = 0xffffffffu, mx = 0u, cnt = 0; double sum = 0;
  for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < n; i += (int)(gridDim.x * blockDim.x)) {
    if (masked && status[i] != 0) continue;
    const float v = fq_value(x, i, mode, center);
    const u32 o = f2ord(v);
    mn = o < mn ? o : mn; mx = o > mx ? o : mx; ++cnt; sum += (double)v;
  }
  mn = c.reduce(mn, MinOp()); mx = c.reduce(mx, MaxOp()); cnt = c.reduce(cnt, SumOp()); sum = c.reduce(sum, SumOp());
  if (c.tid == 0 && cnt) {
    QuantJob* j = &st->qj[slot];
    atomicMax(&j->omin, ~mn); atomicMax(&j->omax, mx); atomicAdd(&j->n, cnt); atomicAdd(&j->sum, sum);
  }
}

__device__ __forceinline__ void fq_range(const QuantJob* j, double* ymin, double* ymax, u64* np) {
  *ymin = (double)ord2f(~j->omin); *ymax = (double)ord2f(j->omax);
  *np = ((*ymax - *ymin) < 0.01) ? 0ull : (u64)((*ymax - *ymin) / 0.01 + 2);
}

__global__ void k_fq_hist(const float* __restrict__ x, const int* __restrict__ status, int masked, int mode, const double* center_p,
                          u32* hist, DevState* st, int slot) {
  __shared__ u32 sh[FQ_SMEM_BINS];
  const int n = st->nb;
  const QuantJob* j = &st->qj[slot];
  if (j->n == 0) return;
  double ymin, ymax; u64 np;
  fq_range(j, &ymin, &ymax, &np);
  if (np == 0) return;
  if (np + 1 > (u64)FQ_BINS_CAP) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(&st->err, (int)ERR_FQ_BINS); return; }
  const double center = center_p ? *center_p : 0.0;
  const bool use_sh = np + 1 <= (u64)FQ_SMEM_BINS;
  if (use_sh) { for (int b = (int)threadIdx.x; b <= (int)np; b += (int)blockDim.x) sh[b] = 0; __syncthreads(); }
  for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < n; i += (int)(gridDim.x * blockDim.x)) {
    if (masked && status[i] != 0) continue;
    const float v = fq_value(x, i, mode, center);
    const double tt = (double)v - ymin;
    double qv = __dadd_rn(__dmul_rn(tt, 100.0), 0.5), fl = floor(qv);      // exact division only next to a bucket edge
    if (qv - fl < 1e-6 || fl + 1.0 - qv < 1e-6) { qv = __dadd_rn(__ddiv_rn(tt, 0.01), 0.5); fl = floor(qv); }
    const u64 b = (u64)fl;
    if (use_sh) atomicAdd(&sh[b], 1u); else atomicAdd(&hist[b], 1u);
  }
  if (use_sh) {
    __syncthreads();
    for (int b = (int)threadIdx.x; b <= (int)np; b += (int)blockDim.x) if (sh[b]) atomicAdd(&hist[b], sh[b]);
  }
}

__global__ void __launch_bounds__(1024) k_fq_pick(u32* hist, DevState* st, int slot) {
  RSI_CTA_SETUP(c);
  QuantJob* j = &st->qj[slot];
  double ymin, ymax; u64 np;
  fq_range(j, &ymin, &ymax, &np);
  const u64 n = j->n;
  double q[3] = {ymin, n ? j->sum / (double)n : 0.0, ymax};
  if (n && np && np + 1 <= (u64)FQ_BINS_CAP) {
    int pick[3], fnz, lnz;
    cta_hist_pick(c, hist, (int)np, n / 4, n / 2, n * 3 / 4, pick, &fnz, &lnz);
    for (int k = 0; k < 3; ++k) if (pick[k] >= 0) q[k] = ymin + (double)pick[k] * 0.01;
    c.sync();
    for (u64 b = (u64)c.tid; b <= np; b += (u64)c.nthr) hist[b] = 0u;
  }
  if (c.tid == 0) { j->ymin = ymin; j->ymax = ymax; j->dy = 0.01; j->np = np; j->q[0] = q[0]; j->q[1] = q[1]; j->q[2] = q[2]; }
}

}  // namespace rsigpu
