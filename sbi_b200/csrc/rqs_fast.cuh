// Spline / activation math of the tensor-core kernels (nsf_tc.cu, nsf_vjp_tc.cu), parameters in
// per-thread registers.  Restates rqs.cuh (itself following nflows 0.14
// rational_quadratic_spline / unconstrained_rational_quadratic_spline,
// oracle/nflows_port/transforms/splines/rational_quadratic.py).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "rqs.cuh"

namespace sbi {
namespace tc {

// ---- epilogue math of the bulk-evaluation path ------------------------------------------------
// Same formulas as the SIMT kernels (rqs.cuh, common.cuh) evaluated with the hardware
// approximations ex2/lg2/rcp (about 2 ulp each) instead of the correctly rounded library calls:
// after the 3xTF32 linears the log-density already carries ~1e-5 of rounding, and these
// functions are 40% of the instructions of this kernel.  tests/test_nsf_tc_gpu.py holds the
// result to the SIMT kernel within 5e-4 and to the fp64 oracle within the common 2e-3.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 1 / (1 + 2^(-x log2 e)); x -> -inf gives rcp(inf) = 0, x -> +inf gives 1
__device__ __forceinline__ float sigmoid_fast(float x) {
  return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x));
}
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : __logf(1.f + __expf(x)); }

// Monotone rational-quadratic spline, forward direction, parameters in registers
// (p[0,K) widths, [K,2K) heights, [2K,3K-1) derivatives; restates rqs_forward of rqs.cuh:
// softmax -> min-size affine -> cumulative knots in [-B,B] -> bin = last knot <= x).
template <int K>
__device__ __forceinline__ void rqs_forward_fast(const float (&p)[32], const RqsConst& c, float x,
                                                 float& y, float& ld) {
  const float B = c.B;
  if (!(x >= -B && x <= B)) { y = x; ld = 0.f; return; }
  float ew[K], eh[K];
  float mw = -INFINITY, mh = -INFINITY;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    ew[i] = p[i] * c.isq;
    eh[i] = p[K + i] * c.isq;
    mw = fmaxf(mw, ew[i]);
    mh = fmaxf(mh, eh[i]);
  }
  float sw = 0.f, sh = 0.f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    ew[i] = __expf(ew[i] - mw);
    eh[i] = __expf(eh[i] - mh);
    sw += ew[i];
    sh += eh[i];
  }
  const float rw = __fdividef(1.f - c.min_w * (float)K, sw);
  const float rh = __fdividef(1.f - c.min_h * (float)K, sh);
  float cw = 0.f, ch = 0.f, lo_w = -B, lo_h = -B;
  float xk = -B, xk1 = B, yk = -B, yk1 = B, r0 = c.edge_raw, r1 = c.edge_raw;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    cw += fmaf(rw, ew[i], c.min_w);
    ch += fmaf(rh, eh[i], c.min_h);
    const float hi_w = (i == K - 1) ? B : fmaf(2.f * B, cw, -B);
    const float hi_h = (i == K - 1) ? B : fmaf(2.f * B, ch, -B);
    if (x >= lo_w) {
      xk = lo_w; xk1 = hi_w; yk = lo_h; yk1 = hi_h;
      r0 = (i == 0) ? c.edge_raw : p[2 * K + (i > 0 ? i - 1 : 0)];
      r1 = (i == K - 1) ? c.edge_raw : p[2 * K + (i < K - 1 ? i : 0)];
    }
    lo_w = hi_w;
    lo_h = hi_h;
  }
  const float wb = xk1 - xk, hb = yk1 - yk;
  const float d0 = c.min_d + softplus_fast(r0), d1 = c.min_d + softplus_fast(r1);
  const float iw = __fdividef(1.f, wb);
  const float delta = hb * iw;
  const float th = (x - xk) * iw;
  const float omt = 1.f - th;
  const float tomt = th * omt;
  const float num = hb * (delta * th * th + d0 * tomt);
  const float den = delta + (d0 + d1 - 2.f * delta) * tomt;
  y = yk + __fdividef(num, den);
  const float dnum = delta * delta * (d1 * th * th + 2.f * delta * tomt + d0 * omt * omt);
  ld = __logf(dnum) - 2.f * __logf(den);
}

// Inverse direction (sampling): x = spline^{-1}(y), ld = log dx/dy; restates rqs_inverse of rqs.cuh
// (bin search on the heights axis, root of the quadratic in the numerically stable form).
template <int K>
__device__ __forceinline__ void rqs_inverse_fast(const float (&p)[32], const RqsConst& c, float yin,
                                                 float& x, float& ld) {
  const float B = c.B;
  if (!(yin >= -B && yin <= B)) { x = yin; ld = 0.f; return; }
  float ew[K], eh[K];
  float mw = -INFINITY, mh = -INFINITY;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    ew[i] = p[i] * c.isq;
    eh[i] = p[K + i] * c.isq;
    mw = fmaxf(mw, ew[i]);
    mh = fmaxf(mh, eh[i]);
  }
  float sw = 0.f, sh = 0.f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    ew[i] = __expf(ew[i] - mw);
    eh[i] = __expf(eh[i] - mh);
    sw += ew[i];
    sh += eh[i];
  }
  const float rw = __fdividef(1.f - c.min_w * (float)K, sw);
  const float rh = __fdividef(1.f - c.min_h * (float)K, sh);
  float cw = 0.f, ch = 0.f, lo_w = -B, lo_h = -B;
  float xk = -B, xk1 = B, yk = -B, yk1 = B, r0 = c.edge_raw, r1 = c.edge_raw;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    cw += fmaf(rw, ew[i], c.min_w);
    ch += fmaf(rh, eh[i], c.min_h);
    const float hi_w = (i == K - 1) ? B : fmaf(2.f * B, cw, -B);
    const float hi_h = (i == K - 1) ? B : fmaf(2.f * B, ch, -B);
    if (yin >= lo_h) {
      xk = lo_w; xk1 = hi_w; yk = lo_h; yk1 = hi_h;
      r0 = (i == 0) ? c.edge_raw : p[2 * K + (i > 0 ? i - 1 : 0)];
      r1 = (i == K - 1) ? c.edge_raw : p[2 * K + (i < K - 1 ? i : 0)];
    }
    lo_w = hi_w;
    lo_h = hi_h;
  }
  const float wb = xk1 - xk, hb = yk1 - yk;
  const float d0 = c.min_d + softplus_fast(r0), d1 = c.min_d + softplus_fast(r1);
  const float delta = __fdividef(hb, wb);
  const float dy = yin - yk;
  const float s2 = d0 + d1 - 2.f * delta;
  const float a = dy * s2 + hb * (delta - d0);
  const float b = hb * d0 - dy * s2;
  const float cc = -delta * dy;
  const float disc = b * b - 4.f * a * cc;
  const float root = __fdividef(2.f * cc, -b - sqrtf(disc));
  x = fmaf(root, wb, xk);
  const float omr = 1.f - root;
  const float tomt = root * omr;
  const float den = delta + s2 * tomt;
  const float dnum = delta * delta * (d1 * root * root + 2.f * delta * tomt + d0 * omr * omr);
  ld = -(__logf(dnum) - 2.f * __logf(den));
}

// Backward of rqs_forward_fast: given gy = dL/dy and gl = dL/d(log-det) returns dL/dx and writes
// dL/dp into g[0, 3K-1) (g[3K-1, 32) = 0).  Same algebra as rqs_backward_reg (rqs.cuh) with the
// hardware approximations of the forward sweep; the bin is located exactly as the forward did.
template <int K>
__device__ __forceinline__ float rqs_backward_fast(const float (&p)[32], const RqsConst& c, float x, float gy,
                                                   float gl, float (&g)[32]) {
  static_assert(3 * K - 1 <= 32, "bins");
#pragma unroll
  for (int i = 0; i < 32; ++i) g[i] = 0.f;
  const float B = c.B;
  if (!(x >= -B && x <= B)) return gy;
  float ew[K], eh[K];
  float mw = -INFINITY, mh = -INFINITY;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    ew[i] = p[i] * c.isq;
    eh[i] = p[K + i] * c.isq;
    mw = fmaxf(mw, ew[i]);
    mh = fmaxf(mh, eh[i]);
  }
  float sw = 0.f, sh = 0.f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    ew[i] = __expf(ew[i] - mw);
    eh[i] = __expf(eh[i] - mh);
    sw += ew[i];
    sh += eh[i];
  }
  const float scw = 1.f - c.min_w * (float)K, sch = 1.f - c.min_h * (float)K;
  const float isw = rcp_approx(sw), ish = rcp_approx(sh);
  const float rw = scw * isw, rh = sch * ish;
  float cw = 0.f, ch = 0.f, lo_w = -B, lo_h = -B;
  float xk = -B, xk1 = B, yk = -B, yk1 = B, r0 = c.edge_raw, r1 = c.edge_raw;
  int b = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    cw += fmaf(rw, ew[i], c.min_w);
    ch += fmaf(rh, eh[i], c.min_h);
    const float hi_w = (i == K - 1) ? B : fmaf(2.f * B, cw, -B);
    const float hi_h = (i == K - 1) ? B : fmaf(2.f * B, ch, -B);
    if (x >= lo_w) {
      b = i;
      xk = lo_w; xk1 = hi_w; yk = lo_h; yk1 = hi_h;
      r0 = (i == 0) ? c.edge_raw : p[2 * K + (i > 0 ? i - 1 : 0)];
      r1 = (i == K - 1) ? c.edge_raw : p[2 * K + (i < K - 1 ? i : 0)];
    }
    lo_w = hi_w;
    lo_h = hi_h;
  }
  const float wb = xk1 - xk, hb = yk1 - yk;
  const float d0 = c.min_d + softplus_fast(r0), d1 = c.min_d + softplus_fast(r1);
  const float iw = rcp_approx(wb);
  const float delta = hb * iw;
  const float th = (x - xk) * iw;
  const float omt = 1.f - th;
  const float tomt = th * omt;
  const float s2 = d0 + d1 - 2.f * delta;
  const float num = hb * (delta * th * th + d0 * tomt);
  const float den = delta + s2 * tomt;
  const float e = d1 * th * th + 2.f * delta * tomt + d0 * omt * omt;
  const float iden = rcp_approx(den), ie = rcp_approx(e);
  const float iden2 = iden * iden;
  const float omt2 = 1.f - 2.f * th;
  const float dnum_dth = hb * (2.f * delta * th + d0 * omt2);
  const float dden_dth = s2 * omt2;
  const float dy_dth = (dnum_dth * den - num * dden_dth) * iden2;
  const float de_dth = 2.f * d1 * th + 2.f * delta * omt2 - 2.f * d0 * omt;
  const float dld_dth = de_dth * ie - 2.f * dden_dth * iden;
  const float dy_ddel = (hb * th * th * den - num * (1.f - 2.f * tomt)) * iden2;
  const float dld_ddel = 2.f * rcp_approx(delta) + 2.f * tomt * ie - 2.f * (1.f - 2.f * tomt) * iden;
  const float dy_dd0 = (hb * tomt * den - num * tomt) * iden2;
  const float dy_dd1 = (-num * tomt) * iden2;
  const float dld_dd0 = omt * omt * ie - 2.f * tomt * iden;
  const float dld_dd1 = th * th * ie - 2.f * tomt * iden;
  const float dy_dhb = (delta * th * th + d0 * tomt) * iden;
  const float Gth = gy * dy_dth + gl * dld_dth;
  const float Gdel = gy * dy_ddel + gl * dld_ddel;
  const float Gd0 = gy * dy_dd0 + gl * dld_dd0;
  const float Gd1 = gy * dy_dd1 + gl * dld_dd1;
  const float gx = Gth * iw;
  const float gxk = -gx;
  const float gwb = -(Gth * th + Gdel * delta) * iw;
  const float ghb = gy * dy_dhb + Gdel * iw;
  const float gA_w = (b >= 1) ? (gxk - gwb) : 0.f;
  const float gB_w = (b <= K - 2) ? gwb : 0.f;
  const float gA_h = (b >= 1) ? (gy - ghb) : 0.f;
  const float gB_h = (b <= K - 2) ? ghb : 0.f;
  const float twoB = 2.f * B;
  {
    // softmax backward of the widths: g_m = p_m (c_m - sum_j p_j c_j) isq,  c_m = d(knots)/d(width_m) weights
    const float sc = scw * twoB;
    float dot = 0.f;
#pragma unroll
    for (int m = 0; m < K; ++m) {
      ew[m] *= isw;
      dot = fmaf(ew[m], sc * ((m < b ? gA_w : 0.f) + (m <= b ? gB_w : 0.f)), dot);
    }
#pragma unroll
    for (int m = 0; m < K; ++m)
      g[m] = ew[m] * (sc * ((m < b ? gA_w : 0.f) + (m <= b ? gB_w : 0.f)) - dot) * c.isq;
  }
  {
    const float sc = sch * twoB;
    float dot = 0.f;
#pragma unroll
    for (int m = 0; m < K; ++m) {
      eh[m] *= ish;
      dot = fmaf(eh[m], sc * ((m < b ? gA_h : 0.f) + (m <= b ? gB_h : 0.f)), dot);
    }
#pragma unroll
    for (int m = 0; m < K; ++m)
      g[K + m] = eh[m] * (sc * ((m < b ? gA_h : 0.f) + (m <= b ? gB_h : 0.f)) - dot) * c.isq;
  }
  // derivatives: d softplus = sigmoid; the two boundary knots carry no parameter
  const float g0 = Gd0 * sigmoid_fast(r0), g1 = Gd1 * sigmoid_fast(r1);
#pragma unroll
  for (int m = 0; m < K - 1; ++m) {
    float v = 0.f;
    if (m == b - 1) v += g0;
    if (m == b) v += g1;
    g[2 * K + m] = v;
  }
  return gx;
}

}  // namespace tc
}  // namespace sbi
