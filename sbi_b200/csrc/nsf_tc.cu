// Tensor-core bulk evaluation and sampling of the neural spline flow: NFlowsFlow.log_prob /
// inverse_transform / sample (/root/reference/sbi/neural_nets/estimators/nflows_flow.py:42-128)
// from 1024 rows.
//
// Same function, same parameter buffer and same evaluation order outside the linears as
// nsf_logprob_kernel (nsf.cu); the ResidualNet linears (nflows ResidualNet, restated in
// oracle/nflows_port/nn/nets/resnet.py; built at flow.py:411-419) run on the 5th-generation
// tensor cores:
//
//   * one CTA = 8 warps (128 rows = 128 TMEM lanes, two threads per row splitting the columns
//     of every epilogue), two CTAs per SM (256 TMEM columns each); the warps take turns issuing
//     the MMAs of a stage (whole warp converged, one elected lane) and the TMA copies of the
//     weight stages (tc_common.cuh);
//   * tcgen05.mma kind::tf32, M = 128; A (activations) is read from TMEM,
//     where the row threads put it with tcgen05.st after splitting every fp32 value into
//     hi = tf32(x) and lo = x - hi; B (weights, pre-split hi/lo and pre-arranged in the K-major
//     no-swizzle UMMA layout by tc_pack_kernel) is streamed by TMA bulk copies into a
//     shared-memory ring; D = A_hi B_hi + A_lo B_hi + A_hi B_lo (3xTF32) accumulates in TMEM in
//     fp32 and comes back with tcgen05.ld for the bias / relu / GLU / spline epilogues;
//   * the context is a K-extension of the hidden operand: A columns are
//     [ hidden (H) | context (C) | 0 ], so the GLU gate W_c ctx is one more small MMA on the
//     same operand and the context never has to be re-staged;
//   * spline, LU (register-resident row against zero-padded 16x16 factors) and base density are
//     per-thread code on the thread's own row; the gate's sigmoid is evaluated while W_1 relu(h)
//     is on the tensor core and the spline of final-layer pass p while pass p+1 is computed;
//   * the same kernel template runs the sampling direction (layers T-1..0, LU^-1, inverse spline).
//
// TMEM columns of a CTA:  [0,64) A_hi | [64,128) A_lo | [128,192) D | [192,256) G (GLU gate);
// the final layer's spline parameters P (32 columns per feature, 2 features per pass)
// alternate between D and G.
#include <cuda_runtime.h>
#include <math.h>
#include <algorithm>
#include <cstdlib>

#include "nsf.cuh"

#include "tc_common.cuh"
#include "nsf_tc_save.cuh"
#include "rqs_fast.cuh"
#include "device.cuh"

namespace sbi {
namespace tc {

// ---- shared memory plan -------------------------------------------------------------------------
struct TcSmem {
  int zs, ctx, lds, lum, bias, bias_stride, ring;   // float offsets
  int bar_bytes, total_bytes;
};
__host__ __device__ inline TcSmem tc_smem_layout(const sbi_nsf_model& m, int stage_cap, int nslot) {
  TcSmem L;
  int fl = 0;
  L.zs = fl;  fl += m.Dp * kRows;
  L.ctx = fl; fl += m.Cp * kRows;
  L.lds = fl; fl += kRows;
  L.lum = fl; fl += 2 * kLuMax * kLuMax + 2 * kLuMax;   // [U 16x16 | L 16x16 | bias 16 | diag 16]
  L.bias_stride = 64 + m.NB * 192 + m.TRmax * 32;
  L.bias = fl; fl += m.T * L.bias_stride;
  fl = (fl + 31) & ~31;
  L.ring = fl; fl += nslot * stage_cap;
  L.bar_bytes = fl * 4;
  L.total_bytes = L.bar_bytes + (nslot + 2) * 8 + 16;
  return L;
}

// Two threads share a row: `half` 0 owns hidden columns [0, HP8/2), `half` 1 owns [HP8/2, HP8)
// (which end in the first context columns).  All epilogues are column-wise, so the
// halves never exchange activations; the spline features of a layer alternate between them.
//
// Per coupling layer the tensor core sees these stages (accumulator barrier in brackets):
//   initial layer -> D [0]
//   per block:  W_c ctx -> G [1],  W_1 relu(h) -> D [0]   (issued together: the gate's sigmoid is
//               evaluated while W_1 runs),  W_2 relu(.) -> D [0]
//   final layer in passes of <= 2 spline features, pass p -> P_(p&1) [p&1]; two passes are in
//               flight, so the spline of pass p runs while pass p+1 is computed.
//
// INV = false: log_prob (logp (R,), optional base-space point `noise` (R,D)).
// INV = true : sampling direction x = T^{-1}(noise | cond): rows.d_input holds the noise, the
//              layers run T-1 .. 0 with LU^{-1} first and the inverse spline; `noise` receives x
//              (R,D) and `logp` (optional) log|det dx/dnoise|  (sbi_b200_nsf_inverse).
//
// SAVE = true (training forward, INV = false): every layer's conditioner intermediates, raw spline
// parameters, layer input and coupling output of the tile go to the activation scratch `save`
// (layout: nsf_tc_save.cuh) for the tensor-core backward kernel (nsf_vjp_tc.cu), together with the
// final base-space point and the row's log-density.
template <int H, int KB, bool INV, bool SAVE = false>
__global__ void __launch_bounds__(kThreads, 2)
nsf_logprob_tc_kernel(const __grid_constant__ sbi_nsf_model m, const __grid_constant__ sbi_nsf_tc tc,
                      const __grid_constant__ sbi_rows rows, float* __restrict__ logp,
                      float* __restrict__ noise, float* __restrict__ save) {
  constexpr int HP8 = (H + 7) & ~7;
  constexpr int NCH = HP8 / 8;      // K-steps / 8-column chunks of the hidden operand
  constexpr int KC0 = H / 8;        // first chunk that holds context columns
  constexpr int NC = HP8 / 2;       // hidden columns per thread (half 0: [0,NC), half 1: [NC,HP8))
  constexpr int NG = NC / 4;        // groups of 4 columns
  constexpr int QC = H - NC;        // first column offset of half 1 that is a context column
  static_assert(HP8 % 8 == 0 && NC % 4 == 0 && H > NC && H <= 64, "hidden width");
  extern __shared__ __align__(128) float sm[];
  const TcSmem L = tc_smem_layout(m, tc.stage_cap, kSlots);
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(sm) + L.bar_bytes);
  uint64_t* bars = full + kSlots;             // two accumulator barriers
  uint32_t* tbase_s = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int C = m.C;
  const int nkc = (H + C + 7) / 8 - KC0;     // K-steps that cover the context columns
  const int64_t ntiles = (rows.R + kRows - 1) / kRows;
  const TcSave SV = tc_save_layout(m.NB, m.TRmax, m.T);

  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) mbar_init(&full[s], 1);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tbase_s)),
                 "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tbase = *tbase_s;

  const float* __restrict__ P = m.d_params;
  float* zs = sm + L.zs;
  float* ctx_s = sm + L.ctx;
  float* lds = sm + L.lds;
  const float* bias_s = sm + L.bias;
  const int half = warp >> 2;                          // which column half of the row
  const int row = ((warp & 3) << 5) | (tid & 31);      // row of the tile = TMEM lane
  const uint32_t tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  const int cbase = half * NC;                          // first hidden column of this thread
  const uint32_t tmine = tlane + cbase;
  RqsConst rc = rqs_const(m);
  rc.K = KB;
  const int D = m.D;

  Issuer iss;
  iss.tbase = __shfl_sync(0xffffffffu, tbase, 0); iss.ring = sm + L.ring; iss.full = full; iss.bars = bars;
  iss.tcw = tc.d_tcw; iss.tab = tc.d_tab; iss.cap = tc.stage_cap; iss.T = m.T;
  iss.it = 0; iss.done = 0; iss.fetched = 0; iss.cov0 = iss.cov1 = 0;
  iss.sbase = 0; iss.lo_off = 0;
  iss.f_tile = blockIdx.x; iss.ntiles = ntiles; iss.tile_step = gridDim.x; iss.f_l = 0; iss.f_s = 0;
  iss.reverse = INV;
  {
    uint32_t el = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    iss.leader = el != 0;
  }
  iss.warp = warp; iss.mine = false;
  iss.pump();     // first kSlots stages
  uint32_t bpar = 0u;          // phase parity of the two accumulator barriers (bit b)

  // all biases of the conditioners, once per CTA (zero beyond the real widths):
  //   per layer [b0 64 | per block: b1 64, b2 64, bc 64 | bf TRmax*32]
  {
    float* bs = sm + L.bias;
    for (int e = tid; e < m.T * L.bias_stride; e += kRowThreads) {
      const int l = e / L.bias_stride, o = e % L.bias_stride;
      const int* LT = m.d_layer_tab + l * SBI_NSF_LAYER_STRIDE;
      float v = 0.f;
      if (o < 64) {
        if (o < H) v = __ldg(P + __ldg(LT + SBI_L_B0) + o);
      } else if (o < 64 + m.NB * 192) {
        const int b = (o - 64) / 192, w = ((o - 64) % 192) / 64, j = (o - 64) % 64;
        if (j < H) v = __ldg(P + __ldg(LT + SBI_L_BLK0 + 6 * b + 1 + 2 * w) + j);
      } else {
        const int q = o - 64 - m.NB * 192, f = q / 32, i = q % 32;
        if (f < __ldg(LT + SBI_L_NTR) && i < 3 * KB - 1) v = __ldg(P + __ldg(LT + SBI_L_BF) + f * m.PR + i);
      }
      bs[e] = v;
    }
  }
  // batch-constant part of the log-density, summed in the order of lu_logdet_total (nsf.cuh)
  float ld_const = 0.f;
  if (half == 1) {
    float tot = 0.f;
    for (int l = 0; l < m.T; ++l) {
      const int* LT = m.d_layer_tab + l * SBI_NSF_LAYER_STRIDE;
      float sacc = 0.f;
      if (__ldg(LT + SBI_L_HAS_LU))
        for (int i = 0; i < D; ++i)
          sacc += logf(softplus_f(__ldg(P + __ldg(LT + SBI_L_LU_DIAG) + i)) + 1e-3f);
      tot += sacc;
    }
    ld_const = INV ? (-tot - m.ld_zscore) : (tot + m.ld_zscore - 0.5f * (float)D * 1.8378770664093453f);
  }

  // operands written / accumulators read: hand TMEM over to the issuing thread
  auto hand_over = [&]() {
    wait_st();
    fence_before();
    group_sync();
  };
  // wait for the accumulators signalled on barrier b
  auto wait_acc = [&](int b) {
    mbar_wait(&bars[b], (bpar >> b) & 1u);
    bpar ^= 1u << b;
    __syncwarp();
    fence_after();
    iss.passed(b);
  };
  // A-operand column j of a hidden layer: activation (j < H), context (H <= j < H+C), zero
  auto acol = [&](int j, float act) -> float {
    const int c = j - H;
    return j < H ? act : ((c < C) ? ctx_s[c * kRows + row] : 0.f);
  };

  // dense LU factors of layer l, zero-padded to 16x16: [U | L | bias 16 | diag 16]
  auto prep_lu = [&](int l) {
    const int* LT = m.d_layer_tab + l * SBI_NSF_LAYER_STRIDE;
    if (!__ldg(LT + SBI_L_HAS_LU)) return;
    const float* lo = P + __ldg(LT + SBI_L_LU_LOWER);
    const float* up = P + __ldg(LT + SBI_L_LU_UPPER);
    const float* dg = P + __ldg(LT + SBI_L_LU_DIAG);
    const float* bi = P + __ldg(LT + SBI_L_LU_BIAS);
    float* U = sm + L.lum;
    float* Lw = U + kLuMax * kLuMax;
    for (int t = tid; t < kLuMax * kLuMax; t += kRowThreads) {
      const int i = t / kLuMax, j = t % kLuMax;
      float u = 0.f, lv = 0.f;
      if (i < D && j < D) {
        if (j > i) u = __ldg(up + i * D - i * (i + 1) / 2 + (j - i - 1));
        else if (j < i) lv = __ldg(lo + i * (i - 1) / 2 + j);
        else u = softplus_f(__ldg(dg + i)) + 1e-3f;
      }
      U[t] = u;
      Lw[t] = lv;
      if (j == 0) Lw[kLuMax * kLuMax + i] = (i < D) ? __ldg(bi + i) : 0.f;
      if (j == i) Lw[kLuMax * kLuMax + kLuMax + i] = (i < D) ? u : 1.f;
    }
  };
  // this thread's NC columns of a hidden-layer A operand; half 1's last columns are context
  auto write_a = [&](const float (&act)[NC]) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = 4 * g + i;
        if (q < QC) a[i] = act[q];
        else a[i] = half ? ((q - QC < C) ? ctx_s[(q - QC) * kRows + row] : 0.f) : act[q];
      }
      store_a4(tlane, cbase + 4 * g, a);
    }
  };
  auto read_acc = [&](int region, float (&d)[NC]) {
#pragma unroll
    for (int g = 0; g < NG; ++g) ld4(tmine + region + 4 * g, d + 4 * g);
    wait_ld();
  };

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * kRows;
    // ---- load + standardise the tile's rows (arithmetic of load_rows, stages.cuh) ----
    {
      const float* st = m.d_stats;
      const int Dp = m.Dp, Cp = m.Cp;
      for (int e = tid; e < kRows * Dp; e += kRowThreads) {
        const int r = e / Dp, d = e % Dp;
        const int64_t gr = row0 + r;
        float val = 0.f;
        if (d < D && gr < rows.R) {
          const int64_t src = rows.d_index ? __ldg(rows.d_index + gr) : gr;
          const float x = __ldg(rows.d_input + src * D + d);
          val = INV ? x : __fadd_rn(__fmul_rn(x, __ldg(st + Dp + d)), __ldg(st + d));
        }
        zs[d * kRows + r] = val;
      }
      for (int e = tid; e < kRows * Cp; e += kRowThreads) {
        const int r = e / Cp, c = e % Cp;
        const int64_t gr = row0 + r;
        float val = 0.f;
        if (c < C && gr < rows.R) {
          const int64_t src = rows.cond_shared ? 0 : (rows.d_index ? __ldg(rows.d_index + gr) : gr);
          val = (__ldg(rows.d_cond + src * C + c) - __ldg(st + 2 * Dp + c)) / __ldg(st + 2 * Dp + Cp + c);
        }
        ctx_s[c * kRows + r] = val;
      }
      if (INV) prep_lu(m.T - 1);
      group_sync();
    }
    // context tail columns [HP8, 64) never change within a tile
    if (half == 0) {
#pragma unroll
      for (int c = NCH; c < 8; ++c) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acol(8 * c + i, 0.f);
        store_a8(tlane, 8 * c, v);
      }
    }
    float ldacc = 0.f;

    for (int li = 0; li < m.T; ++li) {
      SBI_TL(1000 * (li + 1));
      const int l = INV ? m.T - 1 - li : li;
      const NsfLayerView v = layer_view(m, l);
      const int32_t* tab = tc.d_tab + l * SBI_NSF_TC_STRIDE;
      const float* bl = bias_s + l * L.bias_stride;
      const int kid8 = __ldg(tab + 1);
      int stage = 0;
      float h[NC];
      float* svl = SAVE ? save + (size_t)tile * SV.tile_stride + (size_t)l * SV.layer_stride : nullptr;
      if (SAVE && half == 1) tc_save_row16(svl + SV.zin, row, zs, D);      // layer input z_l

      // ---- sampling: z <- U^{-1} L^{-1} (z - b) on the thread's row (order of lu_inverse, nsf.cuh)
      if (INV && half == 1 && __ldg(v.LT + SBI_L_HAS_LU)) {
        const float4* U4 = reinterpret_cast<const float4*>(sm + L.lum);
        const float4* L4 = U4 + kLuMax * kLuMax / 4;
        const float* bias = sm + L.lum + 2 * kLuMax * kLuMax;
        const float* diag = bias + kLuMax;
        float zr[kLuMax];
#pragma unroll
        for (int j = 0; j < kLuMax; ++j) zr[j] = (j < D) ? zs[j * kRows + row] : 0.f;
#pragma unroll
        for (int i = 0; i < kLuMax; ++i) {
          if (i < D) {
            float a = zr[i] - bias[i];
#pragma unroll
            for (int j4 = 0; j4 <= (i - 1) / 4 && i > 0; ++j4) {
              const float4 w = L4[i * (kLuMax / 4) + j4];
              if (4 * j4 + 0 < i) a -= w.x * zr[4 * j4 + 0];
              if (4 * j4 + 1 < i) a -= w.y * zr[4 * j4 + 1];
              if (4 * j4 + 2 < i) a -= w.z * zr[4 * j4 + 2];
              if (4 * j4 + 3 < i) a -= w.w * zr[4 * j4 + 3];
            }
            zr[i] = a;
          }
        }
#pragma unroll
        for (int i = kLuMax - 1; i >= 0; --i) {
          if (i < D) {
            float a = zr[i];
#pragma unroll
            for (int j4 = (i + 1) / 4; j4 < kLuMax / 4; ++j4) {
              const float4 w = U4[i * (kLuMax / 4) + j4];      // padded entries are zero
              if (4 * j4 + 0 > i) a -= w.x * zr[4 * j4 + 0];
              if (4 * j4 + 1 > i) a -= w.y * zr[4 * j4 + 1];
              if (4 * j4 + 2 > i) a -= w.z * zr[4 * j4 + 2];
              if (4 * j4 + 3 > i) a -= w.w * zr[4 * j4 + 3];
            }
            zr[i] = a / diag[i];
            zs[i * kRows + row] = zr[i];
          }
        }
      }
      // ---- initial layer: A = [identity features | 0 ... | context] ----
      // (half 1 ran the LU on this row, so it also writes the identity columns)
      if (half == 1) {
        for (int kk = 0; kk < kid8 / 8; ++kk) {
          float a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = 8 * kk + i;
            a[i] = (j < v.n_id) ? zs[__ldg(v.idf + j) * kRows + row] : 0.f;
          }
          store_a8(tlane, 8 * kk, a);
        }
      } else {
#pragma unroll
        for (int c = KC0; c < NCH; ++c) {
          float a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = acol(8 * c + i, 0.f);
          store_a8(tlane, 8 * c, a);
        }
      }
      hand_over();
      {
        iss.begin(__ldg(tab + 5 + 4 * stage));
        uint32_t acc = 0u;
        iss.block(cD, 0, kid8 / 8, 0, 64, acc);
        iss.block(cD, 8 * KC0, nkc, 64 * kid8, 64, acc);
        iss.end(0);
      }
      ++stage;
      SBI_TL(1000 * (li + 1) + 1);
      // dense LU factors: forward needs this layer's after the spline, sampling needs the next
      // processed layer's before its conditioner; either way the previous contents were last
      // read before the barrier above
      if (!INV) prep_lu(l);
      else if (l > 0) prep_lu(l - 1);
      wait_acc(0);
      const float* blh = bl + cbase;
      {
        float d[NC];
        read_acc(cD, d);
#pragma unroll
        for (int q = 0; q < NC; ++q) h[q] = d[q] + blh[q];     // columns >= H: zero weights + zero bias
      }
      SBI_TL(1000 * (li + 1) + 2);

      // ---- residual blocks ----
      for (int b = 0; b < m.NB; ++b) {
        const float* b1 = blh + 64 + b * 192;
        const float* b2 = b1 + 64;
        const float* bc = b1 + 128;
        // A = [relu(h) | ctx]
        if (SAVE) tc_save_cols<NC>(svl + SV.h(b), row, half, h);
        {
          float a[NC];
#pragma unroll
          for (int q = 0; q < NC; ++q) a[q] = relu_f(h[q]);
          write_a(a);
        }
        hand_over();
        {
          uint32_t accg = 0u;
          iss.begin(__ldg(tab + 5 + 4 * stage));
          iss.block(cG, 8 * KC0, nkc, 0, 64, accg);
          iss.end(1);
          uint32_t acc = 0u;
          iss.begin(__ldg(tab + 5 + 4 * (stage + 1)));
          iss.block(cD, 0, NCH, 0, 64, acc);
          iss.end(0);
        }
        stage += 2;
        SBI_TL(1000 * (li + 1) + 10 * b + 13);
        // gate = sigmoid(Wc ctx + bc) while W1 relu(h) is on the tensor core
        float sg[NC];
        wait_acc(1);
        {
          float g[NC];
          read_acc(cG, g);
#pragma unroll
          for (int q = 0; q < NC; ++q) sg[q] = sigmoid_fast(g[q] + bc[q]);
          if (SAVE) tc_save_cols<NC>(svl + SV.s(b), row, half, sg);
        }
        SBI_TL(1000 * (li + 1) + 10 * b + 14);
        wait_acc(0);
        {
          float d[NC];
          read_acc(cD, d);
#pragma unroll
          for (int q = 0; q < NC; ++q) d[q] = relu_f(d[q] + b1[q]);
          if (SAVE) tc_save_cols<NC>(svl + SV.a1(b), row, half, d);
          write_a(d);
        }
        hand_over();
        {
          uint32_t acc = 0u;
          iss.begin(__ldg(tab + 5 + 4 * stage));
          iss.block(cD, 0, NCH, 0, 64, acc);
          iss.end(0);
        }
        ++stage;
        SBI_TL(1000 * (li + 1) + 10 * b + 15);
        wait_acc(0);
        {
          // h += (W2 a + b2) * gate
          float d[NC];
          read_acc(cD, d);
          if (SAVE) {
#pragma unroll
            for (int q = 0; q < NC; ++q) d[q] += b2[q];
            tc_save_cols<NC>(svl + SV.t2(b), row, half, d);
#pragma unroll
            for (int q = 0; q < NC; ++q) h[q] = fmaf(d[q], sg[q], h[q]);
          } else {
#pragma unroll
            for (int q = 0; q < NC; ++q) h[q] = fmaf(d[q] + b2[q], sg[q], h[q]);
          }
        }
        SBI_TL(1000 * (li + 1) + 10 * b + 16);
      }

      // ---- final layer passes + spline on the transformed features ----
      {
        if (SAVE) tc_save_cols<NC>(svl + SV.hf, row, half, h);
        write_a(h);
        const float* bf = bl + 64 + m.NB * 192;
        const int ns = __ldg(tab);
        const int np = ns - stage;            // passes
        hand_over();
        {
          for (int p = 0; p < 2 && p < np; ++p) {
            uint32_t acc = 0u;
            iss.begin(__ldg(tab + 5 + 4 * (stage + p)));
            iss.block(cD + 64 * p, 0, NCH, 0, __ldg(tab + 6 + 4 * (stage + p)), acc);
            iss.end(p);
          }
        }
        SBI_TL(1000 * (li + 1) + 40);
        for (int p = 0; p < np; ++p) {
          const int aux = __ldg(tab + 7 + 4 * (stage + p));
          const int f0 = aux & 0xffff, nf = aux >> 16;
          wait_acc(p & 1);
          for (int f = 0; f < nf; ++f) {
            if (((f0 + f) & 1) != half) continue;     // warp-uniform: features alternate between halves
            float q[32];
            ld_cols<4>(tlane + cD + 64 * (p & 1) + 32 * f, q);
            wait_ld();
            const float* bff = bf + (f0 + f) * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) q[i] = (i < 3 * KB - 1) ? q[i] + bff[i] : 0.f;
            if (SAVE) tc_save_prm(svl + SV.prm, row, m.TRmax, f0 + f, q);
            const int j = __ldg(v.trf + f0 + f);
            const float x = zs[j * kRows + row];
            float y, ld;
            if (INV) rqs_inverse_fast<KB>(q, rc, x, y, ld);
            else rqs_forward_fast<KB>(q, rc, x, y, ld);
            zs[j * kRows + row] = y;
            ldacc += ld;
          }
          SBI_TL(1000 * (li + 1) + 41 + p);
          if (p + 2 < np) {
            // region p&1 has been read by everyone: pass p+2 may overwrite it
            hand_over();
            {
              uint32_t acc = 0u;
              iss.begin(__ldg(tab + 5 + 4 * (stage + p + 2)));
              iss.block(cD + 64 * (p & 1), 0, NCH, 0, __ldg(tab + 6 + 4 * (stage + p + 2)), acc);
              iss.end(p & 1);
            }
          }
        }
      }

      // ---- LULinear on the row (half 1: it has one spline feature less, and it also writes the
      //      next layer's identity columns):  z <- L (U z) + b, in place ----
      group_sync();      // both halves' spline outputs are in zs
      SBI_TL(1000 * (li + 1) + 50);
      if (SAVE && half == 1) tc_save_row16(svl + SV.v, row, zs, D);         // coupling output v_l
      if (!INV && half == 1 && __ldg(v.LT + SBI_L_HAS_LU)) {
        const float4* U4 = reinterpret_cast<const float4*>(sm + L.lum);
        const float4* L4 = U4 + kLuMax * kLuMax / 4;
        const float* bias = sm + L.lum + 2 * kLuMax * kLuMax;
        float zr[kLuMax];
#pragma unroll
        for (int j = 0; j < kLuMax; ++j) zr[j] = (j < D) ? zs[j * kRows + row] : 0.f;
        // y = U z (upper triangular incl. diagonal; padded entries are zero), same j order as
        // lu_forward (nsf.cuh)
#pragma unroll
        for (int i = 0; i < kLuMax; ++i) {
          if (i < D) {
            float a = 0.f;
#pragma unroll
            for (int j4 = i / 4; j4 < kLuMax / 4; ++j4) {
              const float4 w = U4[i * (kLuMax / 4) + j4];
              if (4 * j4 + 0 >= i) a = fmaf(w.x, zr[4 * j4 + 0], a);
              if (4 * j4 + 1 >= i) a = fmaf(w.y, zr[4 * j4 + 1], a);
              if (4 * j4 + 2 >= i) a = fmaf(w.z, zr[4 * j4 + 2], a);
              if (4 * j4 + 3 >= i) a = fmaf(w.w, zr[4 * j4 + 3], a);
            }
            zr[i] = a;
          }
        }
        // z = L y + b (strictly lower), rows from the bottom so that y_j (j < i) is still intact
#pragma unroll
        for (int i = kLuMax - 1; i >= 0; --i) {
          if (i < D) {
            float a = zr[i];
#pragma unroll
            for (int j4 = 0; j4 <= (i - 1) / 4 && i > 0; ++j4) {
              const float4 w = L4[i * (kLuMax / 4) + j4];
              if (4 * j4 + 0 < i) a = fmaf(w.x, zr[4 * j4 + 0], a);
              if (4 * j4 + 1 < i) a = fmaf(w.y, zr[4 * j4 + 1], a);
              if (4 * j4 + 2 < i) a = fmaf(w.z, zr[4 * j4 + 2], a);
              if (4 * j4 + 3 < i) a = fmaf(w.w, zr[4 * j4 + 3], a);
            }
            zr[i] = a + bias[i];
            zs[i * kRows + row] = zr[i];
          }
        }
      }
    }

    // ---- base density ----
    SBI_TL(9000);
    if (half == 0) lds[row] = ldacc;
    group_sync();
    if (half == 1 && row0 + row < rows.R) {
      if (!INV) {
        float ss = 0.f;
        for (int d = 0; d < D; ++d) ss = fmaf(zs[d * kRows + row], zs[d * kRows + row], ss);
        const float lp = -0.5f * ss + (lds[row] + ldacc) + ld_const;
        if (logp != nullptr) logp[row0 + row] = lp;
        if (SAVE) {
          tc_save_row16(save + (size_t)tile * SV.tile_stride + SV.zt, row, zs, D);
          float* lpt = save + (size_t)tile * SV.tile_stride + SV.lp;
          lpt[row] = lp;
        }
        if (noise != nullptr)
          for (int d = 0; d < D; ++d) noise[(row0 + row) * D + d] = zs[d * kRows + row];
      } else {
        const float* st = m.d_stats;
        for (int d = 0; d < D; ++d)
          noise[(row0 + row) * D + d] = (zs[d * kRows + row] - __ldg(st + d)) / __ldg(st + m.Dp + d);
        if (logp != nullptr) logp[row0 + row] = (lds[row] + ldacc) + ld_const;
      }
    }
    group_sync();   // rows of the next tile are written cooperatively
  }

  fence_before();
  group_sync();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(kCols)
                 : "memory");
}

}  // namespace tc
}  // namespace sbi

// =================================================================================================
// C ABI
// =================================================================================================
using namespace sbi;

static int tc_num_sms() { return sbi::dev_num_sms(); }

// the weight ring has to fit next to a second CTA on the SM
static int tc_plan_slots(const sbi_nsf_model* m, const sbi_nsf_tc* tc) {
  const tc::TcSmem L = tc::tc_smem_layout(*m, tc->stage_cap, tc::kSlots);
  return L.total_bytes <= 112 * 1024 ? tc::kSlots : 0;
}

extern "C" int sbi_b200_nsf_tc_supported(const sbi_nsf_model* m, const sbi_nsf_tc* tc) {
  sbi::DeviceGuard dev_guard_(m ? m->d_params : nullptr);
  if (!m || !tc) return 0;
  if (m->head != SBI_NSF_SPLINE || m->cond_mlp) return 0;   // spline coupling flow with the ResidualNet conditioner
  if (m->H != 50 || m->KB != 10) return 0;        // instantiated hidden width / bin count
  if (m->H + m->C > 64) return 0;                 // context rides in the hidden operand's K range
  if (m->IDp > 48 || m->PR > 32 || m->D > tc::kLuMax) return 0;
  if (m->NB < 1 || m->NB > SBI_NSF_MAX_BLOCKS) return 0;
  if (tc->stage_cap <= 0 || (tc->stage_cap & 31) || tc->n_words <= 0) return 0;
  return tc_plan_slots(m, tc) >= 2 ? 1 : 0;
}

extern "C" int sbi_b200_nsf_tc_pack(const sbi_nsf_model* m, const sbi_nsf_tc* tc, void* stream) {
  sbi::DeviceGuard dev_guard_(m ? m->d_params : nullptr);
  if (!m || !tc || !m->d_params || !tc->d_src || !tc->d_tcw || tc->n_words <= 0) return SBI_EINVAL;
  const int threads = 256, blocks = (tc->n_words + threads - 1) / threads;
  tc::tc_pack_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(m->d_params, tc->d_src,
                                                                       tc->d_tcw, tc->n_words);
  return (int)cudaGetLastError();
}

extern "C" int sbi_b200_nsf_logprob_tc(const sbi_nsf_model* m, const sbi_nsf_tc* tc,
                                       const sbi_rows* rows, float* d_logp, float* d_noise,
                                       void* stream) {
  sbi::DeviceGuard dev_guard_(m ? m->d_params : nullptr);
  if (!m || !tc || !rows || !rows->d_input || !rows->d_cond || rows->R < 0 || !d_logp)
    return SBI_EINVAL;
  if (!tc->d_tab || !tc->d_tcw) return SBI_EINVAL;
  if (!sbi_b200_nsf_tc_supported(m, tc)) return SBI_ESMEM;
  if (rows->R == 0) return 0;
  const int nslot = tc_plan_slots(m, tc);
  const tc::TcSmem L = tc::tc_smem_layout(*m, tc->stage_cap, nslot);
  auto k = tc::nsf_logprob_tc_kernel<50, 10, false>;
  static int smem_set_[sbi::kMaxDev] = {0};
  int& smem_set = smem_set_[sbi::cur_dev()];
  if (smem_set < L.total_bytes) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total_bytes);
    if (e != cudaSuccess) return SBI_ESMEM;
    smem_set = L.total_bytes;
  }
  const int64_t ntiles = (rows->R + tc::kRows - 1) / tc::kRows;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)tc_num_sms() * 2);
  k<<<grid, tc::kThreads, L.total_bytes, (cudaStream_t)stream>>>(*m, *tc, *rows, d_logp, d_noise, nullptr);
  return (int)cudaGetLastError();
}

int sbi::tc::launch_forward_save(const sbi_nsf_model* m, const sbi_nsf_tc* tc, const sbi_rows* rows, float* d_logp,
                                 float* d_save, cudaStream_t s) {
  const int nslot = tc_plan_slots(m, tc);
  const tc::TcSmem L = tc::tc_smem_layout(*m, tc->stage_cap, nslot);
  auto k = tc::nsf_logprob_tc_kernel<50, 10, false, true>;
  static int smem_set_[sbi::kMaxDev] = {0};
  int& smem_set = smem_set_[sbi::cur_dev()];
  if (smem_set < L.total_bytes) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total_bytes) != cudaSuccess)
      return SBI_ESMEM;
    smem_set = L.total_bytes;
  }
  const int grid = (int)((rows->R + tc::kRows - 1) / tc::kRows);      // one tile per CTA: `d_save` slab = blockIdx
  k<<<grid, tc::kThreads, L.total_bytes, s>>>(*m, *tc, *rows, d_logp, nullptr, d_save);
  return (int)cudaGetLastError();
}

extern "C" int sbi_b200_nsf_inverse_tc(const sbi_nsf_model* m, const sbi_nsf_tc* tc,
                                       const sbi_rows* rows, float* d_out, float* d_logabsdet,
                                       void* stream) {
  sbi::DeviceGuard dev_guard_(m ? m->d_params : nullptr);
  if (!m || !tc || !rows || !rows->d_input || !rows->d_cond || rows->R < 0 || !d_out)
    return SBI_EINVAL;
  if (!tc->d_tab || !tc->d_tcw) return SBI_EINVAL;
  if (!sbi_b200_nsf_tc_supported(m, tc)) return SBI_ESMEM;
  if (rows->R == 0) return 0;
  const int nslot = tc_plan_slots(m, tc);
  const tc::TcSmem L = tc::tc_smem_layout(*m, tc->stage_cap, nslot);
  auto k = tc::nsf_logprob_tc_kernel<50, 10, true>;
  static int smem_set_[sbi::kMaxDev] = {0};
  int& smem_set = smem_set_[sbi::cur_dev()];
  if (smem_set < L.total_bytes) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total_bytes);
    if (e != cudaSuccess) return SBI_ESMEM;
    smem_set = L.total_bytes;
  }
  const int64_t ntiles = (rows->R + tc::kRows - 1) / tc::kRows;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)tc_num_sms() * 2);
  k<<<grid, tc::kThreads, L.total_bytes, (cudaStream_t)stream>>>(*m, *tc, *rows, d_logabsdet, d_out, nullptr);
  return (int)cudaGetLastError();
}

#ifdef SBI_TC_TIMELINE
// tuning builds only: copy out and reset the phase timeline of CTA 0; returns the number of (id, clock) pairs
extern "C" int sbi_b200_debug_timeline_fwd(unsigned long long* out, int cap) {
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, sbi::tc::g_tl_n, sizeof(int));
  if (n > cap) n = cap;
  cudaMemcpyFromSymbol(out, sbi::tc::g_tl, (size_t)n * 2 * sizeof(unsigned long long));
  const int zero = 0;
  cudaMemcpyToSymbol(sbi::tc::g_tl_n, &zero, sizeof(int));
  return n;
}
#endif
