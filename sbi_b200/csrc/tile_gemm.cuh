// Row-tile GEMM building blocks (FP32 SIMT, register tiled).
//
// A CTA owns a tile of TM rows (samples).  Every activation lives in shared memory
// feature-major:  act[f][r], r in [0,TM) contiguous, row stride LD = TM + 4 floats
// (keeps 16-byte alignment; LD % 32 == 4 makes consecutive feature rows land on
// different banks).  Weights are streamed in their native PyTorch layout W[n][Kp]
// (out-major, row stride Kp = round4(K), zero padded) by the producer warp.
//
//   forward      Y[n][r]  = sum_k W[n][k] * X[k][r]                 (gemm_fwd_acc)
//   backward-x   dX[k][r] = sum_n W[n][k] * dY[n][r]                (gemm_dx_acc)
//   backward-w   dW[n][k] = sum_r dY[n][r] * X[k][r]  -> global     (gemm_dw)
//
// The native layout serves all three without a transposed copy: forward unrolls k by
// 4 (float4 along a weight row), backward-x walks n with a float4/float2 of k.
#pragma once
#include "common.cuh"

#ifndef SBI_FWD_UNROLL
#define SBI_FWD_UNROLL 2
#endif
#define SBI_PRAGMA(x) _Pragma(#x)
#define SBI_UNROLL(n) SBI_PRAGMA(unroll n)

namespace sbi {

__device__ __forceinline__ float4 ld4(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

template <int TM>
struct Tile {
  static constexpr int LD = TM + 4;
  static constexpr int NRG = TM / 4;                   // thread groups along rows (4 rows each)
  static constexpr int NOG = kConsumerThreads / NRG;   // thread groups along outputs
};

// acc[i][j] += sum_k W[(g + i*ng)][k] * X[k][4*rg + j],  k in [0, 4*K4)
// W rows are `Kp` floats apart, thread's RN output rows are interleaved with stride ng
// (bank-conflict free weight reads across the output groups of one warp).
template <int TM, int RN>
__device__ __forceinline__ void gemm_fwd_acc(float (&acc)[RN][4], const float* __restrict__ X,
                                             int K4, const float* __restrict__ W, int Kp, int g,
                                             int ng, int rg) {
  constexpr int LD = Tile<TM>::LD;
  const float* xp = X + 4 * rg;
  const float* wp = W + (size_t)g * Kp;
  const int wstep = ng * Kp;
  SBI_UNROLL(SBI_FWD_UNROLL)
  for (int k4 = 0; k4 < K4; ++k4) {
    const float4 x0 = ld4(xp + (4 * k4 + 0) * LD);
    const float4 x1 = ld4(xp + (4 * k4 + 1) * LD);
    const float4 x2 = ld4(xp + (4 * k4 + 2) * LD);
    const float4 x3 = ld4(xp + (4 * k4 + 3) * LD);
#pragma unroll
    for (int i = 0; i < RN; ++i) {
      const float4 w = ld4(wp + i * wstep + 4 * k4);
      acc[i][0] = fmaf(w.x, x0.x, acc[i][0]);
      acc[i][1] = fmaf(w.x, x0.y, acc[i][1]);
      acc[i][2] = fmaf(w.x, x0.z, acc[i][2]);
      acc[i][3] = fmaf(w.x, x0.w, acc[i][3]);
      acc[i][0] = fmaf(w.y, x1.x, acc[i][0]);
      acc[i][1] = fmaf(w.y, x1.y, acc[i][1]);
      acc[i][2] = fmaf(w.y, x1.z, acc[i][2]);
      acc[i][3] = fmaf(w.y, x1.w, acc[i][3]);
      acc[i][0] = fmaf(w.z, x2.x, acc[i][0]);
      acc[i][1] = fmaf(w.z, x2.y, acc[i][1]);
      acc[i][2] = fmaf(w.z, x2.z, acc[i][2]);
      acc[i][3] = fmaf(w.z, x2.w, acc[i][3]);
      acc[i][0] = fmaf(w.w, x3.x, acc[i][0]);
      acc[i][1] = fmaf(w.w, x3.y, acc[i][1]);
      acc[i][2] = fmaf(w.w, x3.z, acc[i][2]);
      acc[i][3] = fmaf(w.w, x3.w, acc[i][3]);
    }
  }
}

// Generic forward stage over one weight chunk of `cnt` rows (cnt % 4 == 0):
//   for each thread tile: acc = W_chunk * X ; epi(g, ng, r0, acc)
// epi receives the chunk-local first output row g, the interleave stride ng (output
// row of acc[i] is g + i*ng) and the first tile row r0 = 4*rg.
template <int TM, int RN, class Epi>
__device__ __forceinline__ void gemm_fwd_chunk(const float* __restrict__ X, int K4,
                                               const float* __restrict__ W, int Kp, int cnt,
                                               Epi&& epi) {
  constexpr int NRG = Tile<TM>::NRG, NOG = Tile<TM>::NOG;
  const int rg = threadIdx.x % NRG, og = threadIdx.x / NRG;
  const int ng = cnt / RN;
  for (int g = og; g < ng; g += NOG) {
    float acc[RN][4];
#pragma unroll
    for (int i = 0; i < RN; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    gemm_fwd_acc<TM, RN>(acc, X, K4, W, Kp, g, ng, rg);
    epi(g, ng, 4 * rg, acc);
  }
}

// backward-x over one weight chunk holding rows n in [0,cnt) of W (global rows n0+n):
//   acc[j][c] (k = RK*g + j, r = 4*rg + c) += sum_n W[n][k] * dY[n0+n][r]
template <int TM, int RK, class Epi>
__device__ __forceinline__ void gemm_dx_chunk(const float* __restrict__ dY, int n0, int cnt,
                                              const float* __restrict__ W, int Kp, int Kout,
                                              Epi&& epi) {
  constexpr int LD = Tile<TM>::LD, NRG = Tile<TM>::NRG, NOG = Tile<TM>::NOG;
  static_assert(RK == 2 || RK == 4, "RK");
  const int rg = threadIdx.x % NRG, og = threadIdx.x / NRG;
  const int ngk = (Kout + RK - 1) / RK;   // Kout <= Kp, Kp % 4 == 0
  for (int g = og; g < ngk; g += NOG) {
    float acc[RK][4];
#pragma unroll
    for (int j = 0; j < RK; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const float* dyp = dY + (size_t)n0 * LD + 4 * rg;
    const float* wp = W + RK * g;
#pragma unroll 4
    for (int n = 0; n < cnt; ++n) {
      const float4 dy = ld4(dyp + n * LD);
      float w[RK];
      if (RK == 4) {
        const float4 t = ld4(wp + (size_t)n * Kp);
        w[0] = t.x; w[1] = t.y; w[RK - 2] = t.z; w[RK - 1] = t.w;
      } else {
        const float2 t = *reinterpret_cast<const float2*>(wp + (size_t)n * Kp);
        w[0] = t.x; w[1] = t.y;
      }
#pragma unroll
      for (int j = 0; j < RK; ++j) {
        acc[j][0] = fmaf(w[j], dy.x, acc[j][0]);
        acc[j][1] = fmaf(w[j], dy.y, acc[j][1]);
        acc[j][2] = fmaf(w[j], dy.z, acc[j][2]);
        acc[j][3] = fmaf(w[j], dy.w, acc[j][3]);
      }
    }
    epi(RK * g, 4 * rg, acc);
  }
}

// backward-w:  gW[n*Kp + k] (=|+=) sum_r dY[n][r] * X[k][r],  n < N, k < K
//              gB[n]        (=|+=) sum_r dY[n][r]
// Warp tile 32 n x 16 k; lane (ln = lane%8, lk = lane/8) owns n = nb+ln+8i, k = kb+lk+4j:
// for a fixed i (j) the 8 (4) distinct rows read by a warp sit on distinct banks.
template <int TM>
__device__ __forceinline__ void gemm_dw(const float* __restrict__ dY, int N,
                                        const float* __restrict__ X, int K, int Kp,
                                        float* __restrict__ gW, float* __restrict__ gB,
                                        bool accumulate) {
  constexpr int LD = Tile<TM>::LD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ln = lane & 7, lk = lane >> 3;
  const int tn = (N + 31) / 32, tk = (K + 15) / 16;
  for (int t = warp; t < tn * tk; t += kConsumerThreads / 32) {
    const int nb = (t / tk) * 32, kb = (t % tk) * 16;
    int nrow[4], krow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      nrow[i] = min(nb + ln + 8 * i, N - 1);
      krow[i] = min(kb + lk + 4 * i, K - 1);
    }
    float acc[4][4];
    float sdy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sdy[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
#pragma unroll 2
    for (int r4 = 0; r4 < TM / 4; ++r4) {
      float4 dy[4], x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        dy[i] = ld4(dY + nrow[i] * LD + 4 * r4);
        x[i] = ld4(X + krow[i] * LD + 4 * r4);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sdy[i] += (dy[i].x + dy[i].y) + (dy[i].z + dy[i].w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(dy[i].x, x[j].x, acc[i][j]);
          acc[i][j] = fmaf(dy[i].y, x[j].y, acc[i][j]);
          acc[i][j] = fmaf(dy[i].z, x[j].z, acc[i][j]);
          acc[i][j] = fmaf(dy[i].w, x[j].w, acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = nb + ln + 8 * i;
      if (n >= N) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kb + lk + 4 * j;
        if (k >= K) continue;
        grad_out(gW + (size_t)n * Kp + k, acc[i][j], accumulate);
      }
      if (gB != nullptr && kb == 0 && lk == 0) {
        grad_out(gB + n, sdy[i], accumulate);
      }
    }
  }
}

}  // namespace sbi
