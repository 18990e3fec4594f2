// Monotone rational-quadratic spline (linear tails), one scalar at a time.
//
// Follows the evaluation order of the reference path
//   nflows 0.14 `unconstrained_rational_quadratic_spline` / `rational_quadratic_spline`
//   (restated in oracle/nflows_port/transforms/splines/rational_quadratic.py; call site
//   /root/reference/sbi/neural_nets/net_builders/flow.py:425-432, bin search restated in
//   /root/reference/sbi/utils/torchutils.py:449-463):
//   softmax -> min-width affine -> sequential cumsum -> knots in [-B,B] -> bin = last knot
//   <= x -> gather -> rational quadratic + log|dy/dx|.  Outside [-B,B]: identity, ld = 0.
//
// Parameters for one (row, feature) are read through a stride (they live feature-major in
// shared memory: p[i*st]):  i in [0,K) widths, [K,2K) heights, [2K,3K-1) derivatives.
//
// All functions are __host__ __device__ so that tests/ can exercise exactly this code on
// the CPU (tests/test_device_math_cpu.py compiles it with g++).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SBI_HD __host__ __device__ __forceinline__
#else
#define SBI_HD inline
#endif

namespace sbi {

struct RqsConst {
  int K;          // bins
  float B;        // tail bound
  float isq;      // 1/sqrt(hidden_features): nflows scales widths/heights logits
  float min_w, min_h, min_d;
  float edge_raw; // log(exp(1-min_d)-1): raw derivative at the two boundary knots
};

SBI_HD float rqs_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
SBI_HD float rqs_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
SBI_HD float rqs_mul_add(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(__fmul_rn(a, b), c);   // unfused, like the ATen elementwise ops
#else
  volatile float t = a * b;
  return t + c;
#endif
}

constexpr int kRqsMaxBins = 16;   // bins kept in registers (num_bins <= 16)

struct RqsLoc {
  int b;          // bin
  float klo, khi; // knots b, b+1
};

// softmax numerators of one axis, computed once: e_i = exp(u_i - max u), s = sum e_i
struct RqsAxis {
  float e[kRqsMaxBins];
  float s;
};

SBI_HD void rqs_axis_load(const float* p, int st, int K, float isq, RqsAxis& a) {
  float u[kRqsMaxBins];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < kRqsMaxBins; ++i)
    if (i < K) { u[i] = p[i * st] * isq; m = fmaxf(m, u[i]); }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kRqsMaxBins; ++i)
    if (i < K) { a.e[i] = expf(u[i] - m); s += a.e[i]; }
  a.s = s;
}

// Knots of one normalised axis (widths or heights).  search=true: locate the bin of x;
// search=false: return the knots of bin bsel.
SBI_HD RqsLoc rqs_knots(const RqsAxis& a, int K, float minv, float B, bool search, float x,
                        int bsel) {
  const float scale = 1.f - minv * (float)K;
  float cum = 0.f, lo = -B;
  RqsLoc o;
  o.b = 0; o.klo = -B; o.khi = B;
#pragma unroll
  for (int i = 0; i < kRqsMaxBins; ++i) {
    if (i < K) {
      const float w = minv + scale * (a.e[i] / a.s);
      cum += w;
      const float hi = (i == K - 1) ? B : rqs_mul_add(2.f * B, cum, -B);
      const bool take = search ? (x >= lo) : (i == bsel);
      if (take) { o.b = i; o.klo = lo; o.khi = hi; }
      lo = hi;
    }
  }
  return o;
}

struct RqsBin {
  int b;
  float xk, wb, yk, hb, d0, d1;
  bool inside;
  RqsAxis aw, ah;   // softmax numerators (valid when inside)
};

// REG = true: `p` is a per-thread register array (st == 1); the two derivative reads are then
// done with an unrolled select instead of a dynamic index, so the array never spills.
template <bool REG>
SBI_HD float rqs_pick(const float* pd, int st, int K, int idx) {
  if (!REG) return pd[idx * st];
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < kRqsMaxBins; ++i)
    if (i < K - 1 && i == idx) v = pd[i * st];
  return v;
}

template <bool REG = false>
SBI_HD RqsBin rqs_locate(const float* p, int st, const RqsConst& c, float x, bool inverse) {
  RqsBin o;
  o.inside = (x >= -c.B) && (x <= c.B);
  o.b = 0; o.xk = o.yk = -c.B; o.wb = o.hb = 2.f * c.B; o.d0 = o.d1 = 1.f;
  if (!o.inside) return o;
  const int K = c.K;
  rqs_axis_load(p, st, K, c.isq, o.aw);
  rqs_axis_load(p + K * st, st, K, c.isq, o.ah);
  RqsLoc lw, lh;
  if (!inverse) {
    lw = rqs_knots(o.aw, K, c.min_w, c.B, true, x, 0);
    lh = rqs_knots(o.ah, K, c.min_h, c.B, false, 0.f, lw.b);
    o.b = lw.b;
  } else {
    lh = rqs_knots(o.ah, K, c.min_h, c.B, true, x, 0);
    lw = rqs_knots(o.aw, K, c.min_w, c.B, false, 0.f, lh.b);
    o.b = lh.b;
  }
  o.xk = lw.klo; o.wb = lw.khi - lw.klo;
  o.yk = lh.klo; o.hb = lh.khi - lh.klo;
  const float* pd = p + 2 * K * st;
  const float de = c.min_d + rqs_softplus(c.edge_raw);
  o.d0 = (o.b == 0) ? de : c.min_d + rqs_softplus(rqs_pick<REG>(pd, st, K, o.b - 1));
  o.d1 = (o.b == K - 1) ? de : c.min_d + rqs_softplus(rqs_pick<REG>(pd, st, K, o.b));
  return o;
}

// forward: y = spline(x), ld = log dy/dx
template <bool REG = false>
SBI_HD void rqs_forward(const float* p, int st, const RqsConst& c, float x, float& y, float& ld) {
  const RqsBin q = rqs_locate<REG>(p, st, c, x, false);
  if (!q.inside) { y = x; ld = 0.f; return; }
  const float delta = q.hb / q.wb;
  const float th = (x - q.xk) / q.wb;
  const float tomt = th * (1.f - th);
  const float num = q.hb * (delta * th * th + q.d0 * tomt);
  const float den = delta + (q.d0 + q.d1 - 2.f * delta) * tomt;
  y = q.yk + num / den;
  const float omt = 1.f - th;
  const float dnum = delta * delta * (q.d1 * th * th + 2.f * delta * tomt + q.d0 * omt * omt);
  ld = logf(dnum) - 2.f * logf(den);
}

// inverse: x = spline^{-1}(y), ld = log dx/dy = -log dy/dx
SBI_HD void rqs_inverse(const float* p, int st, const RqsConst& c, float yin, float& x, float& ld) {
  const RqsBin q = rqs_locate(p, st, c, yin, true);
  if (!q.inside) { x = yin; ld = 0.f; return; }
  const float delta = q.hb / q.wb;
  const float dy = yin - q.yk;
  const float s2 = q.d0 + q.d1 - 2.f * delta;
  const float a = dy * s2 + q.hb * (delta - q.d0);
  const float b = q.hb * q.d0 - dy * s2;
  const float cc = -delta * dy;
  const float disc = b * b - 4.f * a * cc;
  const float root = (2.f * cc) / (-b - sqrtf(disc));
  x = root * q.wb + q.xk;
  const float tomt = root * (1.f - root);
  const float den = delta + s2 * tomt;
  const float omr = 1.f - root;
  const float dnum = delta * delta * (q.d1 * root * root + 2.f * delta * tomt + q.d0 * omr * omr);
  ld = -(logf(dnum) - 2.f * logf(den));
}

// backward of rqs_forward for upstream (gy = dL/dy, gl = dL/dld):
//   returns gx = dL/dx and writes dL/dparam into g[i*gst] for all 3K-1 raw params.
SBI_HD float rqs_backward(const float* p, int st, const RqsConst& c, float x, float gy, float gl,
                          float* g, int gst) {
  const int K = c.K;
  const RqsBin q = rqs_locate(p, st, c, x, false);
  if (!q.inside) {
    for (int i = 0; i < 3 * K - 1; ++i) g[i * gst] = 0.f;
    return gy;
  }
  const float wb = q.wb, hb = q.hb, d0 = q.d0, d1 = q.d1;
  const float delta = hb / wb;
  const float th = (x - q.xk) / wb;
  const float omt = 1.f - th;
  const float tomt = th * omt;
  const float s2 = d0 + d1 - 2.f * delta;
  const float num = hb * (delta * th * th + d0 * tomt);
  const float den = delta + s2 * tomt;
  const float e = d1 * th * th + 2.f * delta * tomt + d0 * omt * omt;
  const float iden = 1.f / den, ie = 1.f / e;
  const float omt2 = 1.f - 2.f * th;
  // partials of y and ld w.r.t. (theta, delta, d0, d1, hb-direct)
  const float dnum_dth = hb * (2.f * delta * th + d0 * omt2);
  const float dden_dth = s2 * omt2;
  const float dy_dth = (dnum_dth * den - num * dden_dth) * iden * iden;
  const float de_dth = 2.f * d1 * th + 2.f * delta * omt2 - 2.f * d0 * omt;
  const float dld_dth = de_dth * ie - 2.f * dden_dth * iden;
  const float dy_ddel = (hb * th * th * den - num * (1.f - 2.f * tomt)) * iden * iden;
  const float dld_ddel = 2.f / delta + 2.f * tomt * ie - 2.f * (1.f - 2.f * tomt) * iden;
  const float dy_dd0 = (hb * tomt * den - num * tomt) * iden * iden;
  const float dy_dd1 = (-num * tomt) * iden * iden;
  const float dld_dd0 = omt * omt * ie - 2.f * tomt * iden;
  const float dld_dd1 = th * th * ie - 2.f * tomt * iden;
  const float dy_dhb = (delta * th * th + d0 * tomt) * iden;   // num/hb/den

  const float Gth = gy * dy_dth + gl * dld_dth;
  const float Gdel = gy * dy_ddel + gl * dld_ddel;
  const float Gd0 = gy * dy_dd0 + gl * dld_dd0;
  const float Gd1 = gy * dy_dd1 + gl * dld_dd1;

  const float iw = 1.f / wb;
  const float gx = Gth * iw;
  const float gxk = -Gth * iw;
  const float gwb = -(Gth * th + Gdel * delta) * iw;
  const float ghb = gy * dy_dhb + Gdel * iw;
  const float gyk = gy;

  const int b = q.b;
  // knots: cw[b] (trainable iff 1<=b<=K-1), cw[b+1] (trainable iff b+1<=K-1)
  const float gA_w = (b >= 1) ? (gxk - gwb) : 0.f;      // grad wrt knot b   (widths axis)
  const float gB_w = (b <= K - 2) ? gwb : 0.f;          // grad wrt knot b+1
  const float gA_h = (b >= 1) ? (gyk - ghb) : 0.f;
  const float gB_h = (b <= K - 2) ? ghb : 0.f;
  const float twoB = 2.f * c.B;

  // widths: g_w[m] = 2B*(gA*[m<b] + gB*[m<=b]); softmax backward
  {
    const float scale = (1.f - c.min_w * (float)K) * twoB;
    const float is = 1.f / q.aw.s;
    float dot = 0.f;
#pragma unroll
    for (int m = 0; m < kRqsMaxBins; ++m)
      if (m < K) dot += (q.aw.e[m] * is) * (scale * ((m < b ? gA_w : 0.f) + (m <= b ? gB_w : 0.f)));
#pragma unroll
    for (int m = 0; m < kRqsMaxBins; ++m)
      if (m < K) {
        const float gsm = scale * ((m < b ? gA_w : 0.f) + (m <= b ? gB_w : 0.f));
        g[m * gst] = (q.aw.e[m] * is) * (gsm - dot) * c.isq;
      }
  }
  {
    const float scale = (1.f - c.min_h * (float)K) * twoB;
    const float is = 1.f / q.ah.s;
    float dot = 0.f;
#pragma unroll
    for (int m = 0; m < kRqsMaxBins; ++m)
      if (m < K) dot += (q.ah.e[m] * is) * (scale * ((m < b ? gA_h : 0.f) + (m <= b ? gB_h : 0.f)));
#pragma unroll
    for (int m = 0; m < kRqsMaxBins; ++m)
      if (m < K) {
        const float gsm = scale * ((m < b ? gA_h : 0.f) + (m <= b ? gB_h : 0.f));
        g[(K + m) * gst] = (q.ah.e[m] * is) * (gsm - dot) * c.isq;
      }
  }
  {
    const float* pd = p + 2 * K * st;
    for (int m = 0; m < K - 1; ++m) {
      float v = 0.f;
      if (m == b - 1) v += Gd0 * rqs_sigmoid(pd[m * st]);   // d_b   <- ud[b-1]
      if (m == b) v += Gd1 * rqs_sigmoid(pd[m * st]);       // d_b+1 <- ud[b]
      g[(2 * K + m) * gst] = v;
    }
  }
  return gx;
}

// Same function with the parameters and their gradients in per-thread register arrays (stride 1):
// every loop has a compile-time trip count and every index is a constant after unrolling, so neither
// array is ever addressed dynamically (used by the tensor-core training kernel, nsf_vjp_tc.cu).
template <int K>
SBI_HD float rqs_backward_reg(const float (&p)[32], const RqsConst& c, float x, float gy, float gl,
                              float (&g)[32]) {
  static_assert(3 * K - 1 <= 32 && K <= kRqsMaxBins, "bins");
#pragma unroll
  for (int i = 0; i < 32; ++i) g[i] = 0.f;
  const RqsBin q = rqs_locate<true>(p, 1, c, x, false);
  if (!q.inside) return gy;
  const float wb = q.wb, hb = q.hb, d0 = q.d0, d1 = q.d1;
  const float delta = hb / wb;
  const float th = (x - q.xk) / wb;
  const float omt = 1.f - th;
  const float tomt = th * omt;
  const float s2 = d0 + d1 - 2.f * delta;
  const float num = hb * (delta * th * th + d0 * tomt);
  const float den = delta + s2 * tomt;
  const float e = d1 * th * th + 2.f * delta * tomt + d0 * omt * omt;
  const float iden = 1.f / den, ie = 1.f / e;
  const float omt2 = 1.f - 2.f * th;
  const float dnum_dth = hb * (2.f * delta * th + d0 * omt2);
  const float dden_dth = s2 * omt2;
  const float dy_dth = (dnum_dth * den - num * dden_dth) * iden * iden;
  const float de_dth = 2.f * d1 * th + 2.f * delta * omt2 - 2.f * d0 * omt;
  const float dld_dth = de_dth * ie - 2.f * dden_dth * iden;
  const float dy_ddel = (hb * th * th * den - num * (1.f - 2.f * tomt)) * iden * iden;
  const float dld_ddel = 2.f / delta + 2.f * tomt * ie - 2.f * (1.f - 2.f * tomt) * iden;
  const float dy_dd0 = (hb * tomt * den - num * tomt) * iden * iden;
  const float dy_dd1 = (-num * tomt) * iden * iden;
  const float dld_dd0 = omt * omt * ie - 2.f * tomt * iden;
  const float dld_dd1 = th * th * ie - 2.f * tomt * iden;
  const float dy_dhb = (delta * th * th + d0 * tomt) * iden;
  const float Gth = gy * dy_dth + gl * dld_dth;
  const float Gdel = gy * dy_ddel + gl * dld_ddel;
  const float Gd0 = gy * dy_dd0 + gl * dld_dd0;
  const float Gd1 = gy * dy_dd1 + gl * dld_dd1;
  const float iw = 1.f / wb;
  const float gx = Gth * iw;
  const float gxk = -Gth * iw;
  const float gwb = -(Gth * th + Gdel * delta) * iw;
  const float ghb = gy * dy_dhb + Gdel * iw;
  const float gyk = gy;
  const int b = q.b;
  const float gA_w = (b >= 1) ? (gxk - gwb) : 0.f;
  const float gB_w = (b <= K - 2) ? gwb : 0.f;
  const float gA_h = (b >= 1) ? (gyk - ghb) : 0.f;
  const float gB_h = (b <= K - 2) ? ghb : 0.f;
  const float twoB = 2.f * c.B;
  {
    const float scale = (1.f - c.min_w * (float)K) * twoB;
    const float is = 1.f / q.aw.s;
    float dot = 0.f;
#pragma unroll
    for (int m = 0; m < kRqsMaxBins; ++m)
      if (m < K) dot += (q.aw.e[m] * is) * (scale * ((m < b ? gA_w : 0.f) + (m <= b ? gB_w : 0.f)));
#pragma unroll
    for (int m = 0; m < kRqsMaxBins; ++m)
      if (m < K) {
        const float gsm = scale * ((m < b ? gA_w : 0.f) + (m <= b ? gB_w : 0.f));
        g[m] = (q.aw.e[m] * is) * (gsm - dot) * c.isq;
      }
  }
  {
    const float scale = (1.f - c.min_h * (float)K) * twoB;
    const float is = 1.f / q.ah.s;
    float dot = 0.f;
#pragma unroll
    for (int m = 0; m < kRqsMaxBins; ++m)
      if (m < K) dot += (q.ah.e[m] * is) * (scale * ((m < b ? gA_h : 0.f) + (m <= b ? gB_h : 0.f)));
#pragma unroll
    for (int m = 0; m < kRqsMaxBins; ++m)
      if (m < K && K + m < 32) {
        const float gsm = scale * ((m < b ? gA_h : 0.f) + (m <= b ? gB_h : 0.f));
        g[K + m] = (q.ah.e[m] * is) * (gsm - dot) * c.isq;
      }
  }
#pragma unroll
  for (int m = 0; m < kRqsMaxBins; ++m)
    if (m < K - 1 && 2 * K + m < 32) {
      float v = 0.f;
      if (m == b - 1) v += Gd0 * rqs_sigmoid(p[2 * K + m]);
      if (m == b) v += Gd1 * rqs_sigmoid(p[2 * K + m]);
      g[2 * K + m] = v;
    }
  return gx;
}

}  // namespace sbi
