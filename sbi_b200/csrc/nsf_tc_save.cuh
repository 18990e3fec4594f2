// Activation scratch shared by the tensor-core training kernels: the forward sweep
// (nsf_logprob_tc_kernel<.., SAVE = true>, nsf_tc.cu) writes it, the backward sweep
// (nsf_vjp_tc_kernel, nsf_vjp_tc.cu) reads it.  One slab per 128-row tile:
//
//   per layer l (layer_stride floats):
//     for each residual block b:  h_b | a1_b | t2_b | s_b     each [128 rows][64 columns]
//         h_b  = input of block b (pre-activation)            (relu mask of dW1's dX, X of dW1)
//         a1_b = relu(W1 relu(h_b) + b1)                      (relu mask, X of dW2)
//         t2_b = W2 a1_b + b2        s_b = sigmoid(Wc ctx + bc)   (GLU: block output = t2 * s)
//     hf   = input of the final layer                         [128][64]
//     prm  = raw spline parameters incl. bias                 [128][TRmax][32]
//     zin  = layer input z_l,  v = coupling output (LULinear input)   each [128][16]
//   then per tile:  zt = base-space point z_T [128][16],  lp = log q [128]
//
// Every array of a slab is stored as float4 groups with the ROW index fastest: group g of row r sits at
// float offset (g * 128 + r) * 4, so the 32 rows of a warp write / read 512 contiguous bytes per
// instruction (4 full lines instead of 32 partial sectors).  Groups: 4 adjacent columns of the
// [128][64] arrays (16 groups), parameter quadruple i of feature f (group f * 8 + i) of prm, and
// components 4i..4i+3 of the [128][16] arrays.
// (what the reference keeps as autograd-saved tensors of nflows' ResidualNet / spline transform,
// /root/reference/sbi/neural_nets/net_builders/flow.py:411-432)
#pragma once
#include <cuda_runtime.h>

#include "../../include/sbi_b200.h"

namespace sbi {
namespace tc {

struct TcSave {
  int NB;
  int hf, prm, zin, v;        // float offsets inside a layer slab
  int layer_stride;
  int zt, lp;                 // float offsets inside a tile slab (after the T layer slabs)
  int64_t tile_stride;
  __host__ __device__ int h(int b) const { return (4 * b + 0) * 64 * 128; }
  __host__ __device__ int a1(int b) const { return (4 * b + 1) * 64 * 128; }
  __host__ __device__ int t2(int b) const { return (4 * b + 2) * 64 * 128; }
  __host__ __device__ int s(int b) const { return (4 * b + 3) * 64 * 128; }
};

__host__ __device__ inline TcSave tc_save_layout(int NB, int TRmax, int T) {
  TcSave L;
  L.NB = NB;
  L.hf = 4 * NB * 64 * 128;
  L.prm = L.hf + 64 * 128;
  L.zin = L.prm + TRmax * 32 * 128;
  L.v = L.zin + 16 * 128;
  L.layer_stride = L.v + 16 * 128;
  L.zt = T * L.layer_stride;
  L.lp = L.zt + 16 * 128;
  L.tile_stride = (int64_t)L.lp + 128;
  return L;
}

__device__ __forceinline__ float4* tc_grp(float* q, int g, int row) {
  return reinterpret_cast<float4*>(q) + g * 128 + row;
}
__device__ __forceinline__ const float4* tc_grp(const float* q, int g, int row) {
  return reinterpret_cast<const float4*>(q) + g * 128 + row;
}

// the thread's NC columns (column half `half`: columns [half*NC, half*NC + NC)) of its row; NC % 4 == 0
template <int NC>
__device__ __forceinline__ void tc_save_cols(float* q, int row, int half, const float (&v)[NC]) {
#pragma unroll
  for (int i = 0; i < NC / 4; ++i)
    __stcg(tc_grp(q, half * (NC / 4) + i, row), make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
}
template <int NC>
__device__ __forceinline__ void tc_load_cols(const float* q, int row, int half, float (&v)[NC]) {
#pragma unroll
  for (int i = 0; i < NC / 4; ++i) {
    const float4 t = __ldcg(tc_grp(q, half * (NC / 4) + i, row));
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
// 32 raw spline parameters of feature f of the row
__device__ __forceinline__ void tc_save_prm(float* q, int row, int TRmax, int f, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    __stcg(tc_grp(q, f * 8 + i, row), make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
}
__device__ __forceinline__ void tc_load_prm(const float* q, int row, int TRmax, int f, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = __ldcg(tc_grp(q, f * 8 + i, row));
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
// the row's D <= 16 values of a feature-major shared tile zs[d][128]
__device__ __forceinline__ void tc_save_row16(float* q, int row, const float* zs, int D) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = (4 * i + j < D) ? zs[(4 * i + j) * 128 + row] : 0.f;
    __stcg(tc_grp(q, i, row), make_float4(t[0], t[1], t[2], t[3]));
  }
}
__device__ __forceinline__ void tc_load_row16(const float* q, int row, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldcg(tc_grp(q, i, row));
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
// component j of the row in a [128][16] array
__device__ __forceinline__ float tc_load_row16_at(const float* q, int row, int j) {
  return __ldcg(q + ((j >> 2) * 128 + row) * 4 + (j & 3));
}

// forward sweep of a training step: nsf_logprob_tc_kernel<50, 10, false, SAVE = true>, one tile per CTA
// (defined in nsf_tc.cu; returns a C-ABI status code)
int launch_forward_save(const sbi_nsf_model* m, const sbi_nsf_tc* tc, const sbi_rows* rows, float* d_logp,
                        float* d_save, cudaStream_t s);

}  // namespace tc
}  // namespace sbi
