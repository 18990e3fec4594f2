// Neural-spline-flow kernels: fused log_prob, fused forward+backward (VJP) and inverse
// (sampling) for the flow that sbi's `build_nsf` assembles
// (/root/reference/sbi/neural_nets/net_builders/flow.py:333-460):
//
//   z-score -> T x [ RQ-spline coupling (ResidualNet conditioner, GLU context) -> LULinear ]
//   -> standard normal base.
//
// One CTA = 8 consumer warps + 1 TMA producer warp, owning a tile of TM rows.  All
// activations of the tile stay in shared memory (feature-major), weights are streamed from
// L2 with cp.async.bulk into an mbarrier ring, inputs are read once and one float per row
// is written back.  Both roles run the same stage sequence (template parameter Role), so the
// producer always knows the next chunk without a table.
#pragma once
#include "../../include/sbi_b200.h"
#include "rqs.cuh"
#include "mog.cuh"
#include "stages.cuh"

namespace sbi {

// ---- shared memory plan (row counts are in rows of LD floats) --------------------------
struct NsfSmem {
  int LD;
  // forward
  int U, Z, H, A0, A1, PRM, LDF, Y, Y2, LDACC, LUM;
  // training extras
  int ZS, VS, HS, A1S, T2S, SS, dZ, dU, dH, dT, dG, dPRM, GR, dCTX;
  int ring;        // float offset of the weight ring
  int bar_bytes;   // byte offset of the mbarriers
  int total_bytes;
};

__host__ __device__ inline NsfSmem nsf_smem_layout(const sbi_nsf_model& m, int TM, bool train) {
  NsfSmem L;
  L.LD = TM + 4;
  int rows = 0;
  auto take = [&](int n) { int o = rows * L.LD; rows += n; return o; };
  const int K0p = m.Cp + m.IDp;
  const int prm = m.TRmax * m.PR;   // spline parameters of ALL transformed features
  L.U = take(K0p);
  L.Z = take(m.Dp);
  L.H = take(m.Hp);
  // PRM aliases [A0 | A1 | extra]: the relu/hidden scratch is dead while the final layer runs
  L.A0 = take(m.Hp);
  L.A1 = take(m.Hp);
  L.PRM = L.A0;
  if (prm > 2 * m.Hp) take(prm - 2 * m.Hp);
  L.LDF = take(round4(m.TRmax));
  L.Y = take(m.Dp);
  L.Y2 = take(m.Dp);
  L.LDACC = take(1);
  // dense LU factors of the current layer: [U D*D | L D*D | bias D | diag D] (+1 scalar)
  L.LUM = take((2 * m.D * m.D + 2 * m.D + 4 + L.LD - 1) / L.LD);
  L.ZS = L.VS = L.HS = L.A1S = L.T2S = L.SS = 0;
  L.dZ = L.dU = L.dH = L.dT = L.dG = L.dPRM = L.GR = L.dCTX = 0;
  if (train) {
    L.ZS = take(m.T * m.Dp);
    L.VS = take(m.T * m.Dp);
    L.HS = take((m.NB + 1) * m.Hp);
    L.A1S = take(m.NB * m.Hp);
    L.T2S = take(m.NB * m.Hp);
    L.SS = take(m.NB * m.Hp);
    L.dZ = take(m.Dp);
    L.dU = take(K0p);
    L.dH = take(m.Hp);
    L.dT = take(m.Hp);
    L.dG = take(m.Hp);
    L.dPRM = take(prm);
    L.GR = take(1);
    L.dCTX = take(m.Cp);
  }
  int fl = rows * L.LD;
  fl = (fl + 31) & ~31;   // 128-byte align the ring
  L.ring = fl;
  fl += m.nbuf * m.wcap;
  L.bar_bytes = fl * 4;
  L.total_bytes = L.bar_bytes + 2 * m.nbuf * 8 + 16;
  return L;
}

// ---- per-layer pieces -----------------------------------------------------------------------
struct NsfLayerView {
  const int* LT;    // layer table row
  const int* idf;   // identity feature indices
  const int* trf;   // transformed feature indices
  int n_id, n_tr;
};
__device__ __forceinline__ NsfLayerView layer_view(const sbi_nsf_model& m, int l) {
  NsfLayerView v;
  v.LT = m.d_layer_tab + l * SBI_NSF_LAYER_STRIDE;
  v.n_id = __ldg(v.LT + SBI_L_NID);
  v.n_tr = __ldg(v.LT + SBI_L_NTR);
  v.idf = m.d_feat_tab + __ldg(v.LT + SBI_L_FEAT);
  v.trf = v.idf + v.n_id;
  return v;
}

__device__ __forceinline__ RqsConst rqs_const(const sbi_nsf_model& m) {
  RqsConst c;
  c.K = m.KB; c.B = m.tail_bound; c.isq = m.inv_sqrt_h;
  c.min_w = m.min_bw; c.min_h = m.min_bh; c.min_d = m.min_d; c.edge_raw = m.edge_raw;
  return c;
}

// U[Cp + i] = Zsrc[idf[i]] (identity features feeding the conditioner), pad rows zero
template <int TM>
__device__ __forceinline__ void gather_identity(const sbi_nsf_model& m, const NsfLayerView& v,
                                                const float* Zsrc, float* U) {
  constexpr int LD = Tile<TM>::LD;
  for (int e = threadIdx.x; e < m.IDp * TM; e += kConsumerThreads) {
    const int i = e / TM, r = e % TM;
    U[(m.Cp + i) * LD + r] = (i < v.n_id) ? Zsrc[__ldg(v.idf + i) * LD + r] : 0.f;
  }
  consumer_sync();
}

// Context-only MLP conditioner of the 1-D flow (sbi's ContextSplineMap, flow.py:1419-1478): Linear -> ReLU ->
// NB x [the SAME Linear -> ReLU] ; the final Linear is `final_layer`.  SAVE keeps every layer output
// (HS[0..NB]) for the backward; otherwise two buffers alternate so that the last output lands in L.H.
template <Role R, int TM, int RN, bool SAVE>
__device__ __forceinline__ float* mlp_forward(const sbi_nsf_model& m, const NsfLayerView& v, WPipe& pipe, float* sm,
                                              const NsfSmem& L) {
  constexpr int LD = Tile<TM>::LD;
  const float* __restrict__ P = m.d_params;
  const int Hp = m.Hp, K0p = m.Cp + m.IDp;
  auto buf = [&](int k) { return SAVE ? sm + L.HS + k * Hp * LD : (((m.NB - k) & 1) ? sm + L.A1 : sm + L.H); };
  float* Hout = buf(0);
  {
    const float* b0 = P + __ldg(v.LT + SBI_L_B0);
    fwd_stage<R, TM, RN>(pipe, P + __ldg(v.LT + SBI_L_W0), Hp, K0p, m.rpc0, sm + L.U,
                         [&](int n0, int g, int ng, int r0, float(&acc)[RN][4]) {
#pragma unroll
                           for (int i = 0; i < RN; ++i) {
                             const int n = n0 + g + i * ng;
                             const float b = __ldg(b0 + n);
                             st4(Hout + n * LD + r0, make_float4(relu_f(acc[i][0] + b), relu_f(acc[i][1] + b),
                                                                 relu_f(acc[i][2] + b), relu_f(acc[i][3] + b)));
                           }
                         });
  }
  const float* bh = P + __ldg(v.LT + SBI_L_BLK0 + 1);
  for (int k = 1; k <= m.NB; ++k) {
    const float* Hin = Hout;
    Hout = buf(k);
    fwd_stage<R, TM, RN>(pipe, P + __ldg(v.LT + SBI_L_BLK0), Hp, Hp, m.rpc1, Hin,
                         [&](int n0, int g, int ng, int r0, float(&acc)[RN][4]) {
#pragma unroll
                           for (int i = 0; i < RN; ++i) {
                             const int n = n0 + g + i * ng;
                             const float b = __ldg(bh + n);
                             st4(Hout + n * LD + r0, make_float4(relu_f(acc[i][0] + b), relu_f(acc[i][1] + b),
                                                                 relu_f(acc[i][2] + b), relu_f(acc[i][3] + b)));
                           }
                         });
  }
  return Hout;
}

// ResidualNet conditioner up to the last hidden state (restating nflows ResidualNet,
// oracle/nflows_port/nn/nets/resnet.py).  SAVE keeps every intermediate for the backward.
// Returns the buffer holding the final hidden state.
template <Role R, int TM, int RN, bool SAVE>
__device__ __forceinline__ float* cond_forward(const sbi_nsf_model& m, const NsfLayerView& v,
                                               WPipe& pipe, float* sm, const NsfSmem& L) {
  constexpr int LD = Tile<TM>::LD;
  if (m.cond_mlp) return mlp_forward<R, TM, RN, SAVE>(m, v, pipe, sm, L);
  const float* __restrict__ P = m.d_params;
  const int Hp = m.Hp, Cp = m.Cp, K0p = m.Cp + m.IDp;
  float* U = sm + L.U;
  float* A0 = sm + L.A0;
  float* Hout = SAVE ? sm + L.HS : sm + L.H;
  {
    const float* b0 = P + __ldg(v.LT + SBI_L_B0);
    // MADE adds the context projection's own bias (initial_layer(x) + context_layer(ctx))
    const float* bc0 = m.head == SBI_NSF_MOG ? P + __ldg(v.LT + SBI_L_BC0) : nullptr;
    fwd_stage<R, TM, RN>(pipe, P + __ldg(v.LT + SBI_L_W0), Hp, K0p, m.rpc0, U,
                         [&](int n0, int g, int ng, int r0, float(&acc)[RN][4]) {
#pragma unroll
                           for (int i = 0; i < RN; ++i) {
                             const int n = n0 + g + i * ng;
                             const float b = __ldg(b0 + n) + (bc0 ? __ldg(bc0 + n) : 0.f);
                             const float4 h = make_float4(acc[i][0] + b, acc[i][1] + b,
                                                          acc[i][2] + b, acc[i][3] + b);
                             st4(Hout + n * LD + r0, h);
                             st4(A0 + n * LD + r0, relu4(h));
                           }
                         });
  }
  for (int b = 0; b < m.NB; ++b) {
    const int* BT = v.LT + SBI_L_BLK0 + 6 * b;
    float* Hin = Hout;
    if (SAVE) Hout = Hin + Hp * LD;
    float* A1o = SAVE ? sm + L.A1S + b * Hp * LD : sm + L.A1;
    float* T2o = sm + L.T2S + b * Hp * LD;
    float* So = sm + L.SS + b * Hp * LD;
    const float* b1 = P + __ldg(BT + 1);
    fwd_stage<R, TM, RN>(pipe, P + __ldg(BT + 0), Hp, Hp, m.rpc1, A0,
                         [&](int n0, int g, int ng, int r0, float(&acc)[RN][4]) {
#pragma unroll
                           for (int i = 0; i < RN; ++i) {
                             const int n = n0 + g + i * ng;
                             const float bb = __ldg(b1 + n);
                             st4(A1o + n * LD + r0,
                                 make_float4(relu_f(acc[i][0] + bb), relu_f(acc[i][1] + bb),
                                             relu_f(acc[i][2] + bb), relu_f(acc[i][3] + bb)));
                           }
                         });
    const float* b2 = P + __ldg(BT + 3);
    const float* bc = P + __ldg(BT + 5);
    glu_stage<R, TM, RN>(
        pipe, P + __ldg(BT + 2), Hp, P + __ldg(BT + 4), Cp, Hp, m.rpc2, A1o, U,
        [&](int n0, int g, int ng, int r0, float(&at)[RN][4], float(&ag)[RN][4]) {
#pragma unroll
          for (int i = 0; i < RN; ++i) {
            const int n = n0 + g + i * ng;
            const float bt = __ldg(b2 + n), bg = __ldg(bc + n);
            const float4 hin = ld4(Hin + n * LD + r0);
            float t[4], s[4], h[4];
            const float hi[4] = {hin.x, hin.y, hin.z, hin.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              t[c] = at[i][c] + bt;
              s[c] = sigmoid_f(ag[i][c] + bg);
              h[c] = hi[c] + t[c] * s[c];
            }
            const float4 hv = make_float4(h[0], h[1], h[2], h[3]);
            st4(Hout + n * LD + r0, hv);
            st4(A0 + n * LD + r0, relu4(hv));
            if (SAVE) {
              st4(T2o + n * LD + r0, make_float4(t[0], t[1], t[2], t[3]));
              st4(So + n * LD + r0, make_float4(s[0], s[1], s[2], s[3]));
            }
          }
        });
  }
  return Hout;
}

// final layer: PRM[f*PR + i] = Wf H + bf for ALL transformed features, weights streamed in
// chunks of nf_chunk features
template <Role R, int TM, int RN>
__device__ __forceinline__ void final_layer(const sbi_nsf_model& m, const NsfLayerView& v,
                                            WPipe& pipe, float* sm, const NsfSmem& L,
                                            const float* Hfin) {
  constexpr int LD = Tile<TM>::LD;
  const float* __restrict__ P = m.d_params;
  float* PRM = sm + L.PRM;
  const float* bf = P + __ldg(v.LT + SBI_L_BF);
  fwd_stage<R, TM, RN>(pipe, P + __ldg(v.LT + SBI_L_WF), v.n_tr * m.PR, m.Hp, m.nf_chunk * m.PR, Hfin,
                       [&](int n0, int g, int ng, int r0, float(&acc)[RN][4]) {
#pragma unroll
                         for (int i = 0; i < RN; ++i) {
                           const int n = n0 + g + i * ng;
                           const float b = __ldg(bf + n);
                           st4(PRM + n * LD + r0, make_float4(acc[i][0] + b, acc[i][1] + b,
                                                              acc[i][2] + b, acc[i][3] + b));
                         }
                       });
}

// final layer + spline on every transformed feature (forward or inverse direction)
template <Role R, int TM, int RN, bool INVERSE>
__device__ __forceinline__ void spline_forward(const sbi_nsf_model& m, const NsfLayerView& v,
                                               WPipe& pipe, float* sm, const NsfSmem& L,
                                               const float* Hfin) {
  constexpr int LD = Tile<TM>::LD;
  final_layer<R, TM, RN>(m, v, pipe, sm, L, Hfin);
  if (R == kConsumer && m.head == SBI_NSF_MOG) {
    // mixture-of-Gaussians likelihood of every feature (feature 0 is the wrapper's dummy: no term)
    const float* PRM = sm + L.PRM;
    const float* Z = sm + L.Z;
    float* LDF = sm + L.LDF;
    for (int t = threadIdx.x; t < v.n_tr * TM; t += kConsumerThreads) {
      const int f = t / TM, r = t % TM;
      const int j = __ldg(v.trf + f);
      LDF[f * LD + r] = j == 0 ? 0.f : mog_log_prob(PRM + f * m.PR * LD + r, LD, m.M, m.mog_eps, Z[j * LD + r]);
    }
    consumer_sync();
  } else if (R == kConsumer) {
    const RqsConst rc = rqs_const(m);
    const float* PRM = sm + L.PRM;
    float* Z = sm + L.Z;
    float* LDF = sm + L.LDF;
    for (int t = threadIdx.x; t < v.n_tr * TM; t += kConsumerThreads) {
      const int f = t / TM, r = t % TM;
      const int j = __ldg(v.trf + f);
      const float x = Z[j * LD + r];
      float y, ld;
      if (INVERSE) rqs_inverse(PRM + f * m.PR * LD + r, LD, rc, x, y, ld);
      else rqs_forward(PRM + f * m.PR * LD + r, LD, rc, x, y, ld);
      Z[j * LD + r] = y;
      LDF[f * LD + r] = ld;
    }
    consumer_sync();
  }
}

// LULinear (restating nflows LULinear, oracle/nflows_port/transforms/lu.py).  The dense factors
// of the current layer are built once per layer-tile in shared memory:
//   LUM = [U (D*D, upper incl. diag = softplus(raw)+eps) | L (D*D, strictly lower) | bias D | diag D]
struct LuView {
  const float* U;
  const float* Lw;
  const float* bias;
  const float* diag;
};
__device__ __forceinline__ LuView lu_view(const sbi_nsf_model& m, const float* sm, const NsfSmem& L) {
  LuView w;
  w.U = sm + L.LUM;
  w.Lw = w.U + m.D * m.D;
  w.bias = w.Lw + m.D * m.D;
  w.diag = w.bias + m.D;
  return w;
}
// no barrier inside: callers sync before the factors are read
__device__ __forceinline__ void lu_prepare(const sbi_nsf_model& m, const NsfLayerView& v, float* sm,
                                           const NsfSmem& L) {
  if (!__ldg(v.LT + SBI_L_HAS_LU)) return;
  const float* __restrict__ P = m.d_params;
  const int D = m.D;
  const float* lo = P + __ldg(v.LT + SBI_L_LU_LOWER);
  const float* up = P + __ldg(v.LT + SBI_L_LU_UPPER);
  const float* dg = P + __ldg(v.LT + SBI_L_LU_DIAG);
  const float* bi = P + __ldg(v.LT + SBI_L_LU_BIAS);
  float* U = sm + L.LUM;
  float* Lw = U + D * D;
  for (int t = threadIdx.x; t < D * D; t += kConsumerThreads) {
    const int i = t / D, j = t % D;
    float u = 0.f, l = 0.f;
    if (j > i) u = __ldg(up + i * D - i * (i + 1) / 2 + (j - i - 1));
    else if (j < i) l = __ldg(lo + i * (i - 1) / 2 + j);
    else u = softplus_f(__ldg(dg + i)) + 1e-3f;
    U[t] = u;
    Lw[t] = l;
    if (j == i) {
      Lw[D * D + i] = __ldg(bi + i);        // bias
      Lw[D * D + D + i] = u;                // diag
    }
  }
}

// LDACC[r] += sum_f LDF[f][r]  (this layer's spline log-dets, fixed order over features).
// No barrier: callers have one before LDF is rewritten / LDACC is read.
template <int TM>
__device__ __forceinline__ void fold_ldf(const NsfLayerView& v, float* sm, const NsfSmem& L) {
  constexpr int LD = Tile<TM>::LD;
  for (int r = threadIdx.x; r < TM; r += kConsumerThreads) {
    float a = sm[L.LDACC + r];
    for (int f = 0; f < v.n_tr; ++f) a += sm[L.LDF + f * LD + r];
    sm[L.LDACC + r] = a;
  }
}

// z <- L (U z) + b   (factors from lu_prepare)
template <int TM>
__device__ __forceinline__ void lu_forward(const sbi_nsf_model& m, const NsfLayerView& v,
                                           float* sm, const NsfSmem& L) {
  constexpr int LD = Tile<TM>::LD;
  if (!__ldg(v.LT + SBI_L_HAS_LU)) return;
  const int D = m.D;
  const LuView w = lu_view(m, sm, L);
  float* Z = sm + L.Z;
  float* Y = sm + L.Y;
  for (int t = threadIdx.x; t < D * TM; t += kConsumerThreads) {
    const int i = t / TM, r = t % TM;
    float a = 0.f;
    for (int j = i; j < D; ++j) a = fmaf(w.U[i * D + j], Z[j * LD + r], a);
    Y[i * LD + r] = a;
  }
  consumer_sync();
  for (int t = threadIdx.x; t < D * TM; t += kConsumerThreads) {
    const int i = t / TM, r = t % TM;
    float a = Y[i * LD + r];
    for (int j = 0; j < i; ++j) a = fmaf(w.Lw[i * D + j], Y[j * LD + r], a);
    Z[i * LD + r] = a + w.bias[i];
  }
  consumer_sync();
}

// z <- U^{-1} L^{-1} (z - b)   (sampling direction)
template <int TM>
__device__ __forceinline__ void lu_inverse(const sbi_nsf_model& m, const NsfLayerView& v,
                                           float* sm, const NsfSmem& L) {
  constexpr int LD = Tile<TM>::LD;
  if (!__ldg(v.LT + SBI_L_HAS_LU)) return;
  const int D = m.D;
  const LuView w = lu_view(m, sm, L);
  float* Z = sm + L.Z;
  // one thread per row: forward substitution with unit-lower L, back substitution with U
  for (int r = threadIdx.x; r < TM; r += kConsumerThreads) {
    for (int i = 0; i < D; ++i) {
      float a = Z[i * LD + r] - w.bias[i];
      for (int j = 0; j < i; ++j) a -= w.Lw[i * D + j] * Z[j * LD + r];
      Z[i * LD + r] = a;
    }
    for (int i = D - 1; i >= 0; --i) {
      float a = Z[i * LD + r];
      for (int j = i + 1; j < D; ++j) a -= w.U[i * D + j] * Z[j * LD + r];
      Z[i * LD + r] = a / w.diag[i];
    }
  }
  consumer_sync();
}

// sum over layers of log|det LU| = sum_i log(softplus(raw_i)+eps): identical for every row.
// Cooperative: one (layer, i) term per thread into scratch, then a fixed-order sum.
__device__ __forceinline__ float lu_logdet_total(const sbi_nsf_model& m, float* scratch) {
  const int n = m.T * m.D;
  for (int t = threadIdx.x; t < n; t += kConsumerThreads) {
    const int l = t / m.D, i = t % m.D;
    const int* LT = m.d_layer_tab + l * SBI_NSF_LAYER_STRIDE;
    float val = 0.f;
    if (__ldg(LT + SBI_L_HAS_LU))
      val = logf(softplus_f(__ldg(m.d_params + __ldg(LT + SBI_L_LU_DIAG) + i)) + 1e-3f);
    scratch[t] = val;
  }
  consumer_sync();
  float tot = 0.f;
  for (int l = 0; l < m.T; ++l) {
    float sacc = 0.f;
    for (int i = 0; i < m.D; ++i) sacc += scratch[l * m.D + i];
    tot += sacc;
  }
  consumer_sync();
  return tot;
}

// ---- tile load --------------------------------------------------------------------------------
// Z = zscore(input rows), U[0:C] = standardize(cond rows); zero pads; LDACC = 0
template <int TM>
__device__ __forceinline__ void load_tile(const sbi_nsf_model& m, const sbi_rows& rows,
                                          int64_t row0, float* sm, const NsfSmem& L,
                                          bool raw_input) {
  load_rows<TM>(m.D, m.Dp, m.C, m.Cp, m.d_stats, rows, row0, sm + L.Z, sm + L.U, raw_input);
  for (int r = threadIdx.x; r < TM; r += kConsumerThreads) sm[L.LDACC + r] = 0.f;
  consumer_sync();
}

__device__ __forceinline__ WPipe make_pipe(const sbi_nsf_model& m, float* sm, const NsfSmem& L) {
  return make_pipe(m.nbuf, m.wcap, sm, L.ring, L.bar_bytes);
}

}  // namespace sbi
