// Mixture-of-Gaussians head of sbi's `made` density estimator: per feature the masked network emits
// M x (logit, mean, unconstrained std), interleaved as nflows reshapes them (outputs.reshape(..., M, 3));
// restates MixtureOfGaussiansMADE.log_prob / .sample (oracle/nflows_port/nn/nde/made.py; reference wrapper
// /root/reference/sbi/utils/nn_utils.py:133-201):
//   log MoG(x) = logsumexp_m( log_softmax(logits)_m - 0.5 log 2 pi - log s_m - 0.5 ((x - mu_m) / s_m)^2 ),
//   s_m = softplus(u_m) + eps.
// Parameters are addressed with a stride (feature-major shared-memory tiles), like rqs.cuh.
#pragma once
#include <math.h>

namespace sbi {

constexpr int kMogMax = 16;                  // largest mixture the layouts pack (PR = round4(3M) <= 48)

__device__ __forceinline__ float mog_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// a_m = logit_m - 0.5 (log 2 pi + 2 log s_m + ((x - mu_m) / s_m)^2)
__device__ __forceinline__ float mog_term(const float* p, int st, int m, float eps, float x, float& q, float& is) {
  const float s = mog_softplus(p[(3 * m + 2) * st]) + eps;
  is = 1.f / s;
  q = (x - p[(3 * m + 1) * st]) * is;
  return p[(3 * m) * st] - 0.5f * (1.8378770664093453f + 2.f * logf(s) + q * q);
}

// (runtime loops with recomputation instead of per-thread arrays: this head shares kernels with the
// spline flow, whose register budget it must not disturb)
static __device__ __noinline__ float mog_log_prob(const float* p, int st, int M, float eps, float x) {
  float ml = -INFINITY, ma = -INFINITY, q, is;
  for (int m = 0; m < M; ++m) {
    ml = fmaxf(ml, p[(3 * m) * st]);
    ma = fmaxf(ma, mog_term(p, st, m, eps, x, q, is));
  }
  float sl = 0.f, sa = 0.f;
  for (int m = 0; m < M; ++m) {
    sl += expf(p[(3 * m) * st] - ml);
    sa += expf(mog_term(p, st, m, eps, x, q, is) - ma);
  }
  return (ma + logf(sa)) - (ml + logf(sl));
}

// backward of g * mog_log_prob: writes d/dp into gp (same stride) and returns d/dx
static __device__ __noinline__ float mog_backward(const float* p, int st, int M, float eps, float x, float g, float* gp,
                                           int gst) {
  float ml = -INFINITY, ma = -INFINITY, q, is;
  for (int m = 0; m < M; ++m) {
    ml = fmaxf(ml, p[(3 * m) * st]);
    ma = fmaxf(ma, mog_term(p, st, m, eps, x, q, is));
  }
  float sl = 0.f, sa = 0.f;
  for (int m = 0; m < M; ++m) {
    sl += expf(p[(3 * m) * st] - ml);
    sa += expf(mog_term(p, st, m, eps, x, q, is) - ma);
  }
  float gx = 0.f;
  for (int m = 0; m < M; ++m) {
    const float gam = expf(mog_term(p, st, m, eps, x, q, is) - ma) / sa;
    const float pi = expf(p[(3 * m) * st] - ml) / sl;
    gp[(3 * m) * gst] = g * (gam - pi);
    gp[(3 * m + 1) * gst] = g * gam * q * is;
    const float u = p[(3 * m + 2) * st];
    gp[(3 * m + 2) * gst] = g * gam * (q * q - 1.f) * is * (1.f / (1.f + expf(-u)));
    gx -= g * gam * q * is;
  }
  return gx;
}

// one draw: component by inverse CDF of softmax(logits) at u in [0, 1), then mu + s * n
static __device__ __noinline__ float mog_sample(const float* p, int st, int M, float eps, float u, float n) {
  float ml = -INFINITY;
  for (int m = 0; m < M; ++m) ml = fmaxf(ml, p[(3 * m) * st]);
  float sl = 0.f;
  for (int m = 0; m < M; ++m) sl += expf(p[(3 * m) * st] - ml);
  const float target = u * sl;
  float c = 0.f;
  int pick = M - 1;
  for (int m = 0; m < M; ++m) {
    c += expf(p[(3 * m) * st] - ml);
    if (target < c) { pick = m; break; }
  }
  const float mu = p[(3 * pick + 1) * st];
  const float s = mog_softplus(p[(3 * pick + 2) * st]) + eps;
  return mu + s * n;
}

}  // namespace sbi
