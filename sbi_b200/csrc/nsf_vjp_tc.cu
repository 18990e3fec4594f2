// Tensor-core training step of the neural spline flow: backward sweep of
//   loss = sum_r g_r log q(theta_r | x_r)      (NFlowsFlow.loss + autograd backward,
//   /root/reference/sbi/neural_nets/estimators/nflows_flow.py:99-109,
//   /root/reference/sbi/inference/trainers/base.py:1171-1180)
// for the parameter gradients, on the activations the tensor-core forward sweep
// (nsf_logprob_tc_kernel<..., SAVE = true>, nsf_tc.cu) left in the activation scratch
// (nsf_tc_save.cuh).  Replaces nsf_vjp_kernel (nsf.cu, FP32 SIMT) when only parameter gradients
// are asked for -- the trainer's case.
//
// One CTA = 8 warps owns a tile of 128 rows (row = TMEM lane, two threads per row splitting the
// columns), walks the layers T-1 .. 0 and, per layer, the linears of the conditioner
// (nflows ResidualNet, restated in oracle/nflows_port/nn/nets/resnet.py) in reverse:
//
//   * input-gradient chain  dX = dY W  on tcgen05.mma kind::tf32, M = 128: A = dY from TMEM (written
//     by the row threads, 3xTF32 hi/lo split), B = W^T streamed by TMA bulk copies from the
//     pre-transposed operand blocks (pack.NsfLayout.tc_bwd_plan) through a 2-slot ring; relu / GLU
//     masks, the spline backward (rqs.cuh) and the LULinear backward are per-thread code on the
//     thread's own row between the MMAs;
//   * weight gradients  dW = dY^T X  (K = the 128 rows of the tile) on the same tensor core with BOTH
//     operands from shared memory: the row threads write dY^T and X^T into K-major staging buffers
//     ([row/4][feature][row%4], one padding row per slab so that the 32 rows of a warp hit 32 banks),
//     one M = 64 MMA chain of 16 K-steps per linear (single TF32 pass: the sum over rows averages the
//     operand rounding), accumulators in TMEM, read back by the 16 lanes per warp that hold them and
//     written to this CTA's partial-gradient slab; a ones row appended to X^T yields the bias
//     gradient in the same MMA.  Staging buffers and accumulators are double buffered, so a weight
//     gradient runs on the tensor core while the chain's next epilogue executes.
//
// TMEM columns (512): [0,64) A_hi | [64,128) A_lo | [128,192) D | [192,256) G |
//                     [256,320) A2_hi | [320,384) A2_lo (second A set: spline passes alternate) |
//                     [384,448) dW slot 0 | [448,512) dW slot 1
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

constexpr int kBwdSlots = 2;
constexpr int kBwdCols = 512;
constexpr int cA2 = 256;
constexpr int cW = 384;
constexpr int kStLd = 65;                       // feature rows per K-slab of a staging buffer (64 + 1 pad)
constexpr int kStFloats = 32 * kStLd * 4;       // [128 rows / 4][65][4]

struct BwdSmem {
  int dz, ctx, gr, lum, stg, obuf, ring;   // float offsets
  int bar_bytes, total_bytes;
};
// ldmax: longest packed weight row (floats) of the model, max(Hp, Cp + IDp, Cp)
__host__ __device__ inline BwdSmem bwd_smem_layout(int stage_cap, int ldmax) {
  BwdSmem L;
  int fl = 0;
  L.dz = fl;  fl += 16 * kRows;
  L.ctx = fl; fl += 16 * kRows;
  L.gr = fl;  fl += kRows;
  L.lum = fl; fl += 2 * kLuMax * kLuMax + 2 * kLuMax;
  fl = (fl + 31) & ~31;
  L.stg = fl; fl += 4 * kStFloats;               // [A'0 | B'0 | A'1 | B'1]
  fl = (fl + 31) & ~31;
  L.obuf = fl; fl += 64 * ldmax + 64;            // one weight-gradient block [<= 64 rows][ld] + its bias row
  fl = (fl + 31) & ~31;
  L.ring = fl; fl += kBwdSlots * stage_cap;
  L.bar_bytes = fl * 4;
  L.total_bytes = L.bar_bytes + (kBwdSlots + 2 + 2 + 1) * 8 + 16;
  return L;
}
__host__ __device__ inline int bwd_ldmax(const sbi_nsf_model& m) {
  int a = m.Hp > m.Cp + m.IDp ? m.Hp : m.Cp + m.IDp;
  return a > m.Cp ? a : m.Cp;
}

// where the accumulators of one weight-gradient MMA go in the partial-gradient slab
struct DwGeo {
  int oW, ldw;     // weight block: entry (m, n) at oW + m * ldw + n for n < nX   (ldw % 4 == 0, oW % 4 == 0)
  int oB;          // bias: entry m at oB + m   (column `ones` of the accumulator)
  int nX, ones;    // X columns that are weight columns; index of the ones column
  int Mv, N;       // rows of the block incl. zero padding rows (Mv % 4 == 0); accumulator columns (multiple of 16)
};

// accumulators of one weight-gradient MMA (M = 64: rows 16q .. 16q+15 sit in lanes 0..15 of lane quarter q)
// -> the shared-memory image of the block, laid out exactly like the block in the parameter buffer
// ([Mv][ldw] weights, then the Mv bias entries), from where ONE thread sends it to the CTA's
// partial-gradient slab with two TMA bulk copies (coalesced, asynchronous; scattered per-lane global
// stores of the same data cost 1.3 us per weight gradient).  Columns >= nX of a row are padding and take
// a zero gradient.  One copy of the code for all call sites (the kernel is instruction-fetch bound).
__device__ __noinline__ void dw_read_fn(uint32_t taddr, const DwGeo g, float* obuf, int mrow, int col0,
                                        bool holds_rows) {
  if (col0 >= g.N) return;                          // warp-uniform
  float v[32];
  ld_cols<4>(taddr + col0, v);
  wait_ld();
  if (holds_rows && mrow < g.Mv) {
    float4* orow = reinterpret_cast<float4*>(obuf + mrow * g.ldw + col0);
    float bv = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int n = col0 + 4 * c;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n + i == g.ones) bv = v[4 * c + i];
      if (n < g.ldw)
        orow[c] = make_float4(n < g.nX ? v[4 * c] : 0.f, n + 1 < g.nX ? v[4 * c + 1] : 0.f,
                              n + 2 < g.nX ? v[4 * c + 2] : 0.f, n + 3 < g.nX ? v[4 * c + 3] : 0.f);
    }
    if (g.ones >= col0 && g.ones < col0 + 32) obuf[g.Mv * g.ldw + mrow] = bv;
  }
}

template <int H, int KB>
__global__ void __launch_bounds__(kThreads, 1)
nsf_vjp_tc_kernel(const __grid_constant__ sbi_nsf_model m, const __grid_constant__ sbi_nsf_tc tcb,
                  const __grid_constant__ sbi_rows rows, const float* __restrict__ gout, float g_const,
                  float* __restrict__ gpart, float* __restrict__ loss_acc,
                  const float* __restrict__ save, int accum_first) {
  constexpr int HP8 = (H + 7) & ~7;
  constexpr int NCH = HP8 / 8;
  constexpr int NC = HP8 / 2;
  constexpr int NG = NC / 4;
  static_assert(HP8 % 8 == 0 && NC % 4 == 0 && H > NC && H < HP8 + 1 && HP8 <= 64, "hidden width");
  extern __shared__ __align__(128) float sm[];
  const BwdSmem L = bwd_smem_layout(tcb.stage_cap, bwd_ldmax(m));
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(sm) + L.bar_bytes);
  uint64_t* bars = full + kBwdSlots;          // two accumulator barriers of the dX chain
  uint64_t* dwbar = bars + 2;                 // one barrier per weight-gradient slot
  uint64_t* obar = dwbar + 2;                 // the output image has been read by its bulk copies
  uint32_t* tbase_s = reinterpret_cast<uint32_t*>(obar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = m.C, D = m.D, Cp = m.Cp, Hp = m.Hp;
  const int64_t ntiles = (rows.R + kRows - 1) / kRows;
  const TcSave SV = tc_save_layout(m.NB, m.TRmax, m.T);

  if (tid == 0) {
    for (int s = 0; s < kBwdSlots; ++s) mbar_init(&full[s], 1);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&dwbar[0], 1);
    mbar_init(&dwbar[1], 1);
    mbar_init(obar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tbase_s)),
                 "r"(kBwdCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tbase = *tbase_s;
  SBI_TL(100);

  const float* __restrict__ P = m.d_params;
  float* dzs = sm + L.dz;
  float* ctx_s = sm + L.ctx;
  float* GRs = sm + L.gr;
  float* stg = sm + L.stg;
  const int half = warp >> 2;
  const int row = ((warp & 3) << 5) | lane;
  const uint32_t tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  const int cbase = half * NC;
  const uint32_t tmine = tlane + cbase;
  const int q_one = H - cbase;                 // this thread's column that is X column H (the ones column)
  RqsConst rc = rqs_const(m);
  rc.K = KB;
  float* gp = gpart + (size_t)blockIdx.x * m.n_params;

  IssuerT<kBwdSlots> iss;
  iss.tbase = __shfl_sync(0xffffffffu, tbase, 0); iss.ring = sm + L.ring; iss.full = full; iss.bars = bars;
  iss.tcw = tcb.d_tcw; iss.tab = tcb.d_tab; iss.cap = tcb.stage_cap; iss.T = m.T;
  iss.it = 0; iss.done = 0; iss.fetched = 0; iss.cov0 = iss.cov1 = 0;
  iss.sbase = 0; iss.lo_off = 0;
  iss.f_tile = blockIdx.x; iss.ntiles = ntiles; iss.tile_step = gridDim.x; iss.f_l = 0; iss.f_s = 0;
  iss.reverse = true;
  {
    uint32_t el = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    iss.leader = el != 0;
  }
  iss.warp = warp; iss.mine = false;
  iss.pump();
  uint32_t bpar = 0u;          // phase parity of bars[0], bars[1] (bits 0,1) and dwbar[0], dwbar[1] (bits 2,3)
  uint32_t dwn = 0u;           // weight-gradient MMAs issued so far (slot = dwn & 1, issuing warp = dwn & 7)
  bool pend0 = false, pend1 = false;
  DwGeo geo0, geo1;
  geo0.oW = geo0.ldw = geo0.oB = geo0.nX = geo0.ones = geo0.Mv = geo0.N = 0;
  geo1 = geo0;
  bool accum = accum_first != 0;

  auto hand_over = [&]() {
    wait_st();
    fence_async_smem();
    fence_before();
    group_sync();
  };
  auto wait_acc = [&](int b) {
    mbar_wait(&bars[b], (bpar >> b) & 1u);
    bpar ^= 1u << b;
    __syncwarp();
    fence_after();
    iss.passed(b);
  };
  // element (feature n, row) of a transposed staging buffer
  auto st_put = [&](float* buf, int n, float v) { buf[((row >> 2) * kStLd + n) * 4 + (row & 3)] = v; };
  // accumulators of the weight-gradient MMA in `slot` -> this CTA's partial gradients
  float* obuf = sm + L.obuf;
  bool opend = false;          // the output image holds a block that has not been sent yet
  uint32_t nflush = 0u;        // blocks sent so far (phase of obar)
  DwGeo ogeo = geo0;
  // make `slot` reusable: wait for the MMA chain that last used it and move its accumulators into the
  // output image (whose previous block has been read out by then).  The caller passes a CTA barrier
  // (hand_over / fence_async_smem + group_sync) and then dw_flush() before the next dw_free.
  auto dw_free = [&](int slot) {
    const bool pend = slot ? pend1 : pend0;
    if (!pend) return;
    mbar_wait(&dwbar[slot], (bpar >> (2 + slot)) & 1u);
    bpar ^= 1u << (2 + slot);
    if (nflush > 0u) {
      // the previous block's copies were issued a whole stage ago: they have read the image by now
      if (tid == 0) {
        bulk_wait_read();
        mbar_arrive(obar);
      }
      mbar_wait(obar, (nflush - 1u) & 1u);
    }
    __syncwarp();
    fence_after();
    ogeo = slot ? geo1 : geo0;
    dw_read_fn(tlane + cW + 64 * slot, ogeo, obuf, (warp & 3) * 16 + lane, half * 32, lane < 16);
    fence_before();        // the accumulator reads are ordered before the next MMA into this region
    if (slot) pend1 = false; else pend0 = false;
    opend = true;
  };
  // after the barrier that followed dw_free: one thread sends the image to the partial-gradient slab
  auto dw_flush = [&]() {
    if (!opend) return;
    if (tid == 0) {
      bulk_s2g(gp + ogeo.oW, obuf, (uint32_t)(ogeo.Mv * ogeo.ldw) * 4u, accum);
      bulk_s2g(gp + ogeo.oB, obuf + ogeo.Mv * ogeo.ldw, (uint32_t)ogeo.Mv * 4u, accum);
      bulk_commit();
    }
    opend = false;
    ++nflush;
  };
  // dW = A'^T-staged dY (M = 64 feature rows) x B'-staged X (N feature rows), K = 128 tile rows
  auto dw_issue = [&](int slot, const DwGeo& g) {
    if ((int)(dwn & 7u) == warp) {
      fence_after();
      const uint32_t a_s = __shfl_sync(0xffffffffu, smem_u32(stg + (2 * slot) * kStFloats), 0);
      const uint32_t b_s = __shfl_sync(0xffffffffu, smem_u32(stg + (2 * slot + 1) * kStFloats), 0);
      const int N = __shfl_sync(0xffffffffu, g.N, 0);
      const uint32_t idesc = make_idesc_mn(64, N);
      uint64_t da = make_bdesc(a_s, kStLd * 16u, 128u);
      uint64_t db = make_bdesc(b_s, kStLd * 16u, 128u);
      const uint64_t dstep = (uint64_t)((2u * kStLd * 16u) >> 4);
      const uint32_t d = iss.tbase + cW + 64 * slot;
#pragma unroll 4
      for (int kk = 0; kk < kRows / 8; ++kk) {
        if (iss.leader) mma_tf32_ss(d, da, db, idesc, kk > 0 ? 1u : 0u);
        da += dstep; db += dstep;
      }
      if (iss.leader) commit(&dwbar[slot]);
    }
    if (slot) { pend1 = true; geo1 = g; } else { pend0 = true; geo0 = g; }
    ++dwn;
  };
  auto write_a = [&](const float (&act)[NC], int col0) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = act[4 * g + i];
      store_a4(tlane, col0 + cbase + 4 * g, a);
    }
  };
  auto read_acc = [&](int region, float (&d)[NC]) {
#pragma unroll
    for (int g = 0; g < NG; ++g) ld4(tmine + region + 4 * g, d + 4 * g);
    wait_ld();
  };
  // dense LU factors of layer l, zero-padded to 16x16: [U | L | bias 16 | diag 16]   (as nsf_tc.cu)
  auto prep_lu = [&](int l) {
    const int* LT = m.d_layer_tab + l * SBI_NSF_LAYER_STRIDE;
    if (!__ldg(LT + SBI_L_HAS_LU)) return;
    const float* lo = P + __ldg(LT + SBI_L_LU_LOWER);
    const float* up = P + __ldg(LT + SBI_L_LU_UPPER);
    const float* dg = P + __ldg(LT + SBI_L_LU_DIAG);
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
      if (j == i) Lw[kLuMax * kLuMax + kLuMax + i] = (i < D) ? u : 1.f;
    }
  };

  int iter = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++iter) {
    if (iter > 0) accum = true;
    const int64_t row0 = tile * kRows;
    const float* svt = save + (size_t)tile * SV.tile_stride;
    // ---- per-tile setup: context (standardised like load_rows, stages.cuh), upstream gradient,
    //      loss statistics, d(sum g log q)/dz_T = -g z_T
    {
      const float* st = m.d_stats;
      const int Dp = m.Dp;
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
      const bool live = row0 + row < rows.R;
      const float g = live ? (gout ? __ldg(gout + row0 + row) : g_const) : 0.f;
      if (half == 0) {
        GRs[row] = g;
        float nll = 0.f, bad = 0.f;
        if (live) {
          const float lp = __ldcg(svt + SV.lp + row);
          if (isfinite(lp)) nll = -lp; else bad = 1.f;
        }
        if (loss_acc != nullptr) {
          nll = warp_sum(nll);
          bad = warp_sum(bad);
          if (lane == 0) {
            atomicAdd(loss_acc + 0, nll);
            if (bad != 0.f) atomicAdd(loss_acc + 1, bad);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = __ldcg(tc_grp(svt + SV.zt, i, row));
          // (rows past the end of the batch were never saved: their gradient is exactly zero)
          dzs[(4 * i + 0) * kRows + row] = live ? -g * t.x : 0.f;
          dzs[(4 * i + 1) * kRows + row] = live ? -g * t.y : 0.f;
          dzs[(4 * i + 2) * kRows + row] = live ? -g * t.z : 0.f;
          dzs[(4 * i + 3) * kRows + row] = live ? -g * t.w : 0.f;
        }
      }
      prep_lu(m.T - 1);
      group_sync();
    }
    SBI_TL(101);

    for (int li = 0; li < m.T; ++li) {
      const int l = m.T - 1 - li;
      const NsfLayerView v = layer_view(m, l);
      const int32_t* tab = tcb.d_tab + l * SBI_NSF_TC_STRIDE;
      const float* svl = svt + (size_t)l * SV.layer_stride;
      int stage = 0;
      SBI_TL(1000 * (li + 1));

      // saved activations travel from L2 while the LU section runs: the final-layer input (X of the
      // final layer's weight gradient) and the spline parameters of this thread's first feature
      float xh[NC], qn[32], xn = 0.f;
      int jn = 0;
      tc_load_cols<NC>(svl + SV.hf, row, half, xh);
#pragma unroll
      for (int i = 0; i < 32; ++i) qn[i] = 0.f;
      if (half < v.n_tr) {
        tc_load_prm(svl + SV.prm, row, m.TRmax, half, qn);
        jn = __ldg(v.trf + half);
        xn = tc_load_row16_at(svl + SV.zin, row, jn);
      }

      // ================= LULinear backward (y = U v, z' = L y + b) =================
      if (__ldg(v.LT + SBI_L_HAS_LU)) {
        float vr[kLuMax];
        if (half == 0) tc_load_row16(svl + SV.v, row, vr);      // in flight across the waits below
        // the row-major scratch below lives in the staging buffers of the slot that is reused next
        const int lslot = (int)(dwn & 1u);
        dw_free(lslot);
        SBI_TL(1000 * (li + 1) + 90);
        fence_async_smem();
        group_sync();            // (the dense factors of this layer were built a layer ago: prep_lu below)
        dw_flush();
        SBI_TL(1000 * (li + 1) + 91);
        float* V = stg + (2 * lslot) * kStFloats;
        float* Y = V + 16 * kRows;
        float* DY = Y + 16 * kRows;
        const float* U = sm + L.lum;
        const float* Lw = U + kLuMax * kLuMax;
        float dyr[kLuMax];
        if (half == 0) {
          // y = U v on the thread's row (half 1 does dy meanwhile)
#pragma unroll
          for (int i = 0; i < kLuMax; ++i) {
            float y = 0.f;
#pragma unroll
            for (int j = 0; j < kLuMax; ++j)
              if (j >= i) y = fmaf(U[i * kLuMax + j], vr[j], y);        // padded entries are zero
            if (i < D) {
              V[i * kRows + row] = vr[i];
              Y[i * kRows + row] = y;
            }
          }
        } else {
          // dy = dz + L^T dz (strictly lower part)
          float dzr[kLuMax];
#pragma unroll
          for (int i = 0; i < kLuMax; ++i) dzr[i] = (i < D) ? dzs[i * kRows + row] : 0.f;
#pragma unroll
          for (int i = 0; i < kLuMax; ++i) {
            float dy = dzr[i];
#pragma unroll
            for (int j = 0; j < kLuMax; ++j)
              if (j > i) dy = fmaf(Lw[j * kLuMax + i], dzr[j], dy);
            dyr[i] = dy;
            if (i < D) DY[i * kRows + row] = dy;
          }
        }
        group_sync();
        SBI_TL(1000 * (li + 1) + 92);
        // parameter gradients: one (i,j) pair per thread, reduction over the tile rows (order of
        // lu_backward, nsf.cu)
        {
          const int o_lo = __ldg(v.LT + SBI_L_LU_LOWER), o_up = __ldg(v.LT + SBI_L_LU_UPPER);
          const int o_dg = __ldg(v.LT + SBI_L_LU_DIAG), o_bi = __ldg(v.LT + SBI_L_LU_BIAS);
          for (int t = tid; t < D * D + D; t += kRowThreads) {
            float a = 0.f;
            float* dst;
            // dot products over the 128 tile rows, four independent partial sums; every lane reads a
            // different feature row (same bank at the same offset), so lane k starts 4k rows in
            const int rot = 4 * lane;
            auto dot = [&](const float* p, const float* q) {
              float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
              for (int r = 0; r < kRows; r += 4) {
                const int rr = (r + rot) & (kRows - 1);
                const float4 x4 = *reinterpret_cast<const float4*>(p + rr);
                const float4 y4 = *reinterpret_cast<const float4*>(q + rr);
                a0 = fmaf(x4.x, y4.x, a0); a1 = fmaf(x4.y, y4.y, a1);
                a2 = fmaf(x4.z, y4.z, a2); a3 = fmaf(x4.w, y4.w, a3);
              }
              return (a0 + a1) + (a2 + a3);
            };
            auto rsum = [&](const float* p) {
              float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
              for (int r = 0; r < kRows; r += 4) {
                const float4 x4 = *reinterpret_cast<const float4*>(p + ((r + rot) & (kRows - 1)));
                a0 += x4.x; a1 += x4.y; a2 += x4.z; a3 += x4.w;
              }
              return (a0 + a1) + (a2 + a3);
            };
            if (t < D * D) {
              const int i = t / D, j = t % D;
              if (i > j) {
                a = dot(dzs + i * kRows, Y + j * kRows);
                dst = gp + o_lo + i * (i - 1) / 2 + j;
              } else if (i < j) {
                a = dot(DY + i * kRows, V + j * kRows);
                dst = gp + o_up + i * D - i * (i + 1) / 2 + (j - i - 1);
              } else {
                a = dot(DY + i * kRows, V + i * kRows);
                const float gs = rsum(GRs);
                a = (a + gs / U[i * kLuMax + i]) * sigmoid_f(__ldg(P + o_dg + i));
                dst = gp + o_dg + i;
              }
            } else {
              const int i = t - D * D;
              a = rsum(dzs + i * kRows);
              dst = gp + o_bi + i;
            }
            *dst = accum ? (*dst + a) : a;
          }
          // the arrays are padded to a multiple of 4 entries: padding takes a zero gradient (every
          // entry of the slab is written by this kernel; nothing is zero-filled beforehand)
          if (!accum && tid >= kRowThreads - 4) {
            const int k = tid - (kRowThreads - 4);
            const int ntri = D * (D - 1) / 2;
            const int o = k == 0 ? o_lo : k == 1 ? o_up : k == 2 ? o_dg : o_bi;
            const int n = k < 2 ? ntri : D;
            for (int e = n; e < ((n + 3) & ~3); ++e) gp[o + e] = 0.f;
          }
        }
        group_sync();
        SBI_TL(1000 * (li + 1) + 93);
        if (half == 1) {
          // dv = U^T dy
#pragma unroll
          for (int j = 0; j < kLuMax; ++j) {
            if (j < D) {
              float a = 0.f;
#pragma unroll
              for (int i = 0; i < kLuMax; ++i)
                if (i <= j) a = fmaf(U[i * kLuMax + j], dyr[i], a);
              dzs[j * kRows + row] = a;
            }
          }
        }
        group_sync();
      }
      if (l > 0) prep_lu(l - 1);     // next layer's dense factors: first read a whole layer of barriers later
      SBI_TL(1000 * (li + 1) + 1);

      // ================= final layer + spline backward, passes of <= 2 features =================
      const int np = __ldg(tab + 1);
      const int oWF = __ldg(v.LT + SBI_L_WF), oBF = __ldg(v.LT + SBI_L_BF);
      {
        for (int p = 0; p < np; ++p) {
          const int f = 2 * p + half;
          const int nf = min(2, v.n_tr - 2 * p);
          const bool has = f < v.n_tr;                     // warp-uniform
          const int aset = (p & 1) * cA2;
          // this pass's parameters were requested a pass (or the LU section) ago; request the next
          float q[32];
          const float x = xn;
          const int j = jn;
#pragma unroll
          for (int i = 0; i < 32; ++i) q[i] = qn[i];
          if (f + 2 < v.n_tr) {
            tc_load_prm(svl + SV.prm, row, m.TRmax, f + 2, qn);
            jn = __ldg(v.trf + f + 2);
            xn = tc_load_row16_at(svl + SV.zin, row, jn);
          }
          if (p >= 2) wait_acc(p & 1);                     // the MMA that read this A set is done
          const int slot = (int)(dwn & 1u);
          dw_free(slot);
          float* As = stg + (2 * slot) * kStFloats;
          float* Bs = As + kStFloats;
          if (has) {
            float dq[32];
            const float gx = rqs_backward_fast<KB>(q, rc, x, dzs[j * kRows + row], GRs[row], dq);
            dzs[j * kRows + row] = gx;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float a[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) a[i] = dq[8 * c + i];
              store_a8(tlane, aset + 32 * half + 8 * c, a);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) st_put(As, 32 * half + i, dq[i]);
          }
          if (p < 2) {      // the X operand (final-layer input) of pass p - 2 is still in this slot
#pragma unroll
            for (int q = 0; q < NC; ++q) st_put(Bs, cbase + q, q == q_one ? 1.f : xh[q]);
          }
          hand_over();
          {
            uint32_t acc = p > 0 ? 1u : 0u;
            iss.begin(__ldg(tab + 5 + 4 * (stage + p)));
            iss.block(cD, aset, 4 * nf, 0, 64, acc);
            iss.end(p & 1);
          }
          DwGeo g;
          g.oW = oWF + 2 * p * m.PR * Hp; g.ldw = Hp; g.oB = oBF + 2 * p * m.PR;
          g.nX = H; g.ones = H; g.Mv = 32 * nf; g.N = 64;
          dw_issue(slot, g);
        dw_flush();
          dw_flush();
          SBI_TL(1000 * (li + 1) + 10 + p);
        }
        stage += np;
      }
      // (loads of saved activations are requested one wait ahead of their use throughout the blocks)
      float sv[NC], t2[NC];
      if (m.NB > 0) {
        tc_load_cols<NC>(svl + SV.s(m.NB - 1), row, half, sv);
        tc_load_cols<NC>(svl + SV.t2(m.NB - 1), row, half, t2);
      }
      for (int p = max(0, np - 2); p < np; ++p) wait_acc(p & 1);
      float dh[NC];
      read_acc(cD, dh);
      SBI_TL(1000 * (li + 1) + 20);

      // ================= residual blocks, last to first =================
      for (int b = m.NB - 1; b >= 0; --b) {
        const int* BT = v.LT + SBI_L_BLK0 + 6 * b;
        float dT[NC], a1[NC];
        tc_load_cols<NC>(svl + SV.a1(b), row, half, a1);
        {
          // ---- dWc = dG^T ctx  (GLU gate; no input gradient wanted for the context)
          const int slot = (int)(dwn & 1u);
          dw_free(slot);
          float* As = stg + (2 * slot) * kStFloats;
          float* Bs = As + kStFloats;
          const int Nc = (Cp + 1 + 15) & ~15;
#pragma unroll
          for (int q = 0; q < NC; ++q) {
            const float dhq = dh[q], s = sv[q];
            dT[q] = dhq * s;
            st_put(As, cbase + q, dhq * t2[q] * s * (1.f - s));
          }
          for (int n = half * (Nc / 2); n < (half + 1) * (Nc / 2); ++n)
            st_put(Bs, n, n < C ? ctx_s[n * kRows + row] : (n == Cp ? 1.f : 0.f));
          fence_async_smem();
          group_sync();
          DwGeo g;
          g.oW = __ldg(BT + 4); g.ldw = Cp; g.oB = __ldg(BT + 5); g.nX = Cp; g.ones = Cp; g.Mv = Hp; g.N = Nc;
          dw_issue(slot, g);
        dw_flush();
          dw_flush();
        }
        SBI_TL(1000 * (li + 1) + 30 + 10 * b);
        // ---- dW2 = dT^T a1 ;  dA1 = (dT W2) * [a1 > 0]
        {
          const int slot = (int)(dwn & 1u);
          dw_free(slot);
          SBI_TL(1000 * (li + 1) + 70 + 10 * b);
          float* As = stg + (2 * slot) * kStFloats;
          float* Bs = As + kStFloats;
#pragma unroll
          for (int q = 0; q < NC; ++q) {
            st_put(As, cbase + q, dT[q]);
            st_put(Bs, cbase + q, q == q_one ? 1.f : a1[q]);
          }
          SBI_TL(1000 * (li + 1) + 71 + 10 * b);
          write_a(dT, 0);
          SBI_TL(1000 * (li + 1) + 72 + 10 * b);
          hand_over();
          SBI_TL(1000 * (li + 1) + 73 + 10 * b);
          {
            uint32_t acc = 0u;
            iss.begin(__ldg(tab + 5 + 4 * stage));
            iss.block(cD, 0, NCH, 0, 64, acc);
            iss.end(0);
          }
          SBI_TL(1000 * (li + 1) + 74 + 10 * b);
          ++stage;
          DwGeo g;
          g.oW = __ldg(BT + 2); g.ldw = Hp; g.oB = __ldg(BT + 3); g.nX = H; g.ones = H; g.Mv = Hp; g.N = 64;
          dw_issue(slot, g);
        dw_flush();
          dw_flush();
        }
        SBI_TL(1000 * (li + 1) + 31 + 10 * b);
        float hb[NC];
        tc_load_cols<NC>(svl + SV.h(b), row, half, hb);
        wait_acc(0);
        float dA[NC];
        read_acc(cD, dA);
#pragma unroll
        for (int q = 0; q < NC; ++q) dA[q] = a1[q] > 0.f ? dA[q] : 0.f;
        SBI_TL(1000 * (li + 1) + 32 + 10 * b);
        // ---- dW1 = dA1^T relu(h_b) ;  dh += (dA1 W1) * [h_b > 0]
        {
          const int slot = (int)(dwn & 1u);
          dw_free(slot);
          float* As = stg + (2 * slot) * kStFloats;
          float* Bs = As + kStFloats;
#pragma unroll
          for (int q = 0; q < NC; ++q) {
            st_put(As, cbase + q, dA[q]);
            st_put(Bs, cbase + q, q == q_one ? 1.f : relu_f(hb[q]));
          }
          write_a(dA, 0);
          hand_over();
          {
            uint32_t acc = 0u;
            iss.begin(__ldg(tab + 5 + 4 * stage));
            iss.block(cG, 0, NCH, 0, 64, acc);
            iss.end(1);
          }
          ++stage;
          DwGeo g;
          g.oW = __ldg(BT + 0); g.ldw = Hp; g.oB = __ldg(BT + 1); g.nX = H; g.ones = H; g.Mv = Hp; g.N = 64;
          dw_issue(slot, g);
        dw_flush();
          dw_flush();
        }
        SBI_TL(1000 * (li + 1) + 33 + 10 * b);
        if (b > 0) {
          tc_load_cols<NC>(svl + SV.s(b - 1), row, half, sv);
          tc_load_cols<NC>(svl + SV.t2(b - 1), row, half, t2);
        }
        wait_acc(1);
        {
          float d[NC];
          read_acc(cG, d);
#pragma unroll
          for (int q = 0; q < NC; ++q) dh[q] += hb[q] > 0.f ? d[q] : 0.f;
        }
        SBI_TL(1000 * (li + 1) + 34 + 10 * b);
      }

      // ================= initial layer: dW0 = dh^T [ctx | id | 1], d(identity features) =================
      {
        const int K0p = Cp + m.IDp;
        const int N0 = (K0p + 1 + 15) & ~15;
        const int slot = (int)(dwn & 1u);
        dw_free(slot);
        float* As = stg + (2 * slot) * kStFloats;
        float* Bs = As + kStFloats;
#pragma unroll
        for (int q = 0; q < NC; ++q) st_put(As, cbase + q, dh[q]);
        for (int n = half * (N0 / 2); n < (half + 1) * (N0 / 2); ++n) {
          float val = 0.f;
          if (n < Cp) val = n < C ? ctx_s[n * kRows + row] : 0.f;
          else if (n < K0p) {
            const int i = n - Cp;
            if (i < v.n_id) val = tc_load_row16_at(svl + SV.zin, row, __ldg(v.idf + i));
          } else if (n == K0p) val = 1.f;
          st_put(Bs, n, val);
        }
        write_a(dh, 0);
        hand_over();
        {
          uint32_t acc = 0u;
          iss.begin(__ldg(tab + 5 + 4 * stage));
          iss.block(cD, 0, NCH, 0, 16, acc);
          iss.end(0);
        }
        ++stage;
        DwGeo g;
        g.oW = __ldg(v.LT + SBI_L_W0); g.ldw = K0p; g.oB = __ldg(v.LT + SBI_L_B0); g.nX = K0p; g.ones = K0p;
        g.Mv = Hp; g.N = N0;
        dw_issue(slot, g);
        dw_flush();
        SBI_TL(1000 * (li + 1) + 60);
        wait_acc(0);
        if (half == 0) {
          float d[16];
          ld8(tlane + cD, d);
          ld8(tlane + cD + 8, d + 8);
          wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < v.n_id) dzs[__ldg(v.idf + j) * kRows + row] += d[j];
        }
        fence_before();
        group_sync();
      }
    }
    SBI_TL(9000);
    // the tile's last weight gradients
    for (int k = 0; k < 2; ++k) {
      dw_free((int)((dwn + k) & 1u));     // older slot first
      fence_async_smem();
      group_sync();
      dw_flush();
    }
  }
  SBI_TL(9001);

  if (tid == 0) bulk_wait_all();
  fence_before();
  group_sync();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(kBwdCols)
                 : "memory");
}

}  // namespace tc
}  // namespace sbi

// =================================================================================================
// C ABI
// =================================================================================================
using namespace sbi;

static int vjp_tc_ok(const sbi_nsf_model* m, const sbi_nsf_tc* tcf, const sbi_nsf_tc* tcb) {
  if (!sbi_b200_nsf_tc_supported(m, tcf)) return 0;
  if (!tcb || !tcb->d_tab || !tcb->d_tcw || tcb->stage_cap <= 0 || (tcb->stage_cap & 31)) return 0;
  if (m->IDp > 16 || m->Cp + m->IDp + 1 > 64 || m->Cp + 1 > 64) return 0;
  const tc::BwdSmem L = tc::bwd_smem_layout(tcb->stage_cap, tc::bwd_ldmax(*m));
  return L.total_bytes <= 227 * 1024 ? 1 : 0;
}

extern "C" int sbi_b200_nsf_vjp_tc_supported(const sbi_nsf_model* m, const sbi_nsf_tc* tc_fwd,
                                             const sbi_nsf_tc* tc_bwd) {
  if (!m || !tc_fwd || !tc_bwd) return 0;
  return vjp_tc_ok(m, tc_fwd, tc_bwd);
}

// rows of one forward + backward launch pair (one tile per CTA, every SM busy once)
static int64_t vjp_tc_chunk_rows() { return (int64_t)tc::kRows * sbi::dev_num_sms(); }

extern "C" int sbi_b200_nsf_vjp_tc_parts(int64_t R) {
  const int64_t ntiles = (R + tc::kRows - 1) / tc::kRows;
  return (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, sbi::dev_num_sms()));
}

extern "C" int64_t sbi_b200_nsf_vjp_tc_save_bytes(const sbi_nsf_model* m, int64_t R) {
  if (!m || R < 1) return 0;
  const tc::TcSave SV = tc::tc_save_layout(m->NB, m->TRmax, m->T);
  return (int64_t)sizeof(float) * SV.tile_stride * sbi_b200_nsf_vjp_tc_parts(R);
}

extern "C" int sbi_b200_nsf_vjp_tc(const sbi_nsf_model* m, const sbi_nsf_tc* tc_fwd, const sbi_nsf_tc* tc_bwd,
                                   const sbi_rows* rows, const float* d_gout, float g_const, float* d_logp,
                                   float* d_gpart, float* d_loss_acc, float* d_save, int64_t save_bytes,
                                   void* stream) {
  sbi::DeviceGuard dev_guard_(m ? m->d_params : nullptr);
  if (!m || !tc_fwd || !tc_bwd || !rows || !rows->d_input || !rows->d_cond || rows->R < 1 || !d_gpart ||
      !d_save)
    return SBI_EINVAL;
  if (!vjp_tc_ok(m, tc_fwd, tc_bwd)) return SBI_ESMEM;
  if (save_bytes < sbi_b200_nsf_vjp_tc_save_bytes(m, rows->R)) return SBI_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const tc::BwdSmem Lb = tc::bwd_smem_layout(tc_bwd->stage_cap, tc::bwd_ldmax(*m));
  auto kb = tc::nsf_vjp_tc_kernel<50, 10>;
  {
    static int set_b_[sbi::kMaxDev] = {0};
    int& set_b = set_b_[sbi::cur_dev()];
    if (set_b < Lb.total_bytes) {
      if (cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, Lb.total_bytes) != cudaSuccess)
        return SBI_ESMEM;
      set_b = Lb.total_bytes;
    }
  }
  // chunks of one tile per SM: forward sweep (saves activations) then backward sweep of the same rows;
  // later chunks accumulate into the partial-gradient slabs
  const int64_t chunk = vjp_tc_chunk_rows();
  for (int64_t r0 = 0; r0 < rows->R; r0 += chunk) {
    sbi_rows rr = *rows;
    rr.R = std::min<int64_t>(chunk, rows->R - r0);
    if (rows->d_index) rr.d_index = rows->d_index + r0;
    else {
      rr.d_input = rows->d_input + r0 * m->D;
      if (!rows->cond_shared) rr.d_cond = rows->d_cond + r0 * m->C;
    }
    const int grid = (int)((rr.R + tc::kRows - 1) / tc::kRows);
    const int rc = tc::launch_forward_save(m, tc_fwd, &rr, d_logp ? d_logp + r0 : nullptr, d_save, s);
    if (rc) return rc;
    kb<<<grid, tc::kThreads, Lb.total_bytes, s>>>(*m, *tc_bwd, rr, d_gout ? d_gout + r0 : nullptr, g_const,
                                                  d_gpart, d_loss_acc, d_save, r0 > 0 ? 1 : 0);
  }
  return (int)cudaGetLastError();
}

#ifdef SBI_TC_TIMELINE
// tuning builds only: copy out and reset the phase timeline of CTA 0; returns the number of (id, clock) pairs
extern "C" int sbi_b200_debug_timeline_bwd(unsigned long long* out, int cap) {
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
