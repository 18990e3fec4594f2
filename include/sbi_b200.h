/* sbi_b200 -- C ABI of the B200-native hot path of sbi (density-estimator training and
 * posterior evaluation).  Plain pointers and sizes only; no torch types.
 *
 * Every pointer named d_* is a DEVICE pointer (sm_100a), every h_* a HOST pointer.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 * All functions return 0 on success, a negative SBI_E* code on argument errors, or a
 * positive cudaError_t value if a CUDA call failed.  Nothing here synchronises unless the
 * name ends in _host (those take host buffers and block until the result is on the host).
 *
 * Reference interface replaced (file:line under /root/reference):
 *   sbi/neural_nets/estimators/nflows_flow.py:77-97   NFlowsFlow.log_prob   -> sbi_b200_nsf_logprob
 *   sbi/neural_nets/estimators/nflows_flow.py:99-109  NFlowsFlow.loss + autograd backward
 *                                                      (trainers/base.py:1171-1187)  -> sbi_b200_nsf_vjp
 *   sbi/neural_nets/estimators/nflows_flow.py:111-128 NFlowsFlow.sample     -> sbi_b200_nsf_inverse
 *   sbi/neural_nets/estimators/nflows_flow.py:42-75   inverse_transform     -> sbi_b200_nsf_logprob (z_out)
 *   sbi/inference/trainers/base.py:1181-1187          clip_grad_norm_ + Adam.step -> sbi_b200_reduce_partials,
 *                                                                                  sbi_b200_adam_clip_step
 */
#ifndef SBI_B200_H
#define SBI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBI_B200_ABI_VERSION 1

#define SBI_EINVAL (-1)   /* bad argument */
#define SBI_ESMEM (-2)    /* model does not fit the shared-memory budget of one CTA */
#define SBI_ENOGPU (-3)   /* no sm_100 device */

/* per-layer descriptor table: SBI_NSF_LAYER_STRIDE ints per coupling layer */
#define SBI_NSF_LAYER_STRIDE 64
#define SBI_NSF_MAX_BLOCKS 8
enum {
  SBI_L_NID = 0,      /* number of identity (conditioner-input) features */
  SBI_L_NTR = 1,      /* number of transformed features */
  SBI_L_W0 = 2,       /* float offset: initial layer  [Hp][Cp+IDp], columns = [ctx | id] */
  SBI_L_B0 = 3,
  SBI_L_WF = 4,       /* final layer [n_tr*PR][Hp]; feature f owns rows f*PR .. f*PR+3K-2 */
  SBI_L_BF = 5,
  SBI_L_LU_LOWER = 6, /* nflows LULinear params, native shapes */
  SBI_L_LU_UPPER = 7,
  SBI_L_LU_DIAG = 8,
  SBI_L_LU_BIAS = 9,
  SBI_L_FEAT = 10,    /* offset into feat_tab: n_id identity features then n_tr transformed */
  SBI_L_HAS_LU = 11,
  SBI_L_BC0 = 12,     /* head == SBI_NSF_MOG: second bias of the initial layer (MADE's context_layer.bias) */
  SBI_L_BLK0 = 16     /* per residual block b, 6 ints at SBI_L_BLK0+6*b: W1,B1,W2,B2,WC,BC */
};
/* what the conditioner's outputs parameterise (`head`):
 *   SBI_NSF_SPLINE (0): rational-quadratic coupling transform + LULinear per layer, N(0, I) base (`nsf`)
 *   SBI_NSF_MOG    (1): `made`: ONE masked residual conditioner (nflows MADE, masks folded into the packed
 *                       weights) whose outputs are, per feature, M x (logit, mean, unconstrained std) of a
 *                       mixture of Gaussians; log q = sum_f log MoG_f(z_f), no base density
 *                       (sbi/neural_nets/net_builders/flow.py:37-112, sbi/utils/nn_utils.py:133-201:
 *                       MADEMoGWrapper prepends a dummy feature, so D here = features + 1 and feature 0
 *                       -- whose input the caller sets to 0 -- is excluded from the likelihood). */
#define SBI_NSF_SPLINE 0
#define SBI_NSF_MOG 1

/* Neural spline flow (sbi `posterior_nn("nsf")` / `likelihood_nn("nsf")`,
 * reference builder sbi/neural_nets/net_builders/flow.py:333-460). */
typedef struct {
  int32_t D, C, H, NB, KB, T;      /* input dim, context dim, hidden, res-blocks, bins, layers */
  int32_t Dp, Cp, IDp, Hp, PR;     /* padded: round4(D), round4(C), round4(max n_id), round4(H), round4(3KB-1) */
  int32_t TRmax;                   /* max transformed features over layers */
  int32_t nf_chunk;                /* features per final-layer weight chunk */
  int32_t rpc0, rpc1, rpc2;        /* rows per weight chunk: initial, hidden, hidden+context (GLU) */
  int32_t wcap, nbuf;              /* weight ring: floats per slot, slots */
  int32_t n_params;                /* floats in d_params */
  float tail_bound, inv_sqrt_h, min_bw, min_bh, min_d, edge_raw;
  int32_t head, M;                 /* SBI_NSF_SPLINE / SBI_NSF_MOG; mixture components (PR = round4(3M)) */
  int32_t cond_mlp;                /* 1: the conditioner is the context-only MLP of the 1-D flow (flow.py:401-408,
                                    * ContextSplineMap :1419-1478): relu(W0 ctx + b0), then NB applications of ONE
                                    * shared hidden layer (W at SBI_L_BLK0, bias at +1) with relu, then the final layer */
  float mog_eps;                   /* std = softplus(.) + mog_eps */
  float ld_zscore;                 /* sum_d log|scale_d| of the input z-score transform */
  const float* d_params;           /* packed parameters (see sbi_b200/pack.py) */
  const int32_t* d_layer_tab;      /* T * SBI_NSF_LAYER_STRIDE */
  const int32_t* d_feat_tab;
  const float* d_stats;            /* [shift(Dp) | scale(Dp) | ctx_mean(Cp) | ctx_std(Cp)] */
} sbi_nsf_model;

/* Optional outputs / inputs of the row kernels; any pointer may be NULL. */
typedef struct {
  const float* d_input;            /* (R, D) row-major, or the gather source when d_index != NULL */
  const float* d_cond;             /* (R, C) row-major, (1, C) when cond_shared, or gather source */
  const int64_t* d_index;          /* (R,) row indices into d_input/d_cond (device-resident data set) */
  int64_t R;
  int32_t cond_shared;             /* 1: one condition row for all R rows (posterior at x_o) */
} sbi_rows;

int sbi_b200_abi_version(void);
int sbi_b200_device_ok(void);      /* 1 if device 0 is sm_100, else 0 */

/* log q(input | cond) for R rows.  d_logp (R,) ; d_noise (R, D) optional (the base-space
 * point z, i.e. NFlowsFlow.inverse_transform). */
int sbi_b200_nsf_logprob(const sbi_nsf_model* m, const sbi_rows* rows, float* d_logp,
                         float* d_noise, void* stream);

/* Vector-Jacobian product of sum_r g_r * log q_r: forward + backward in one kernel.
 *   d_gout (R,) upstream gradient per row, or NULL with g_const used for every row.
 *   d_logp (R,) optional forward output.
 *   d_gpart  (n_part, n_params) per-CTA partial parameter gradients (written, not
 *            accumulated); n_part = sbi_b200_nsf_vjp_parts(R).
 *   d_ginput (R, D), d_gcond (R, C) optional input / condition gradients.
 *   d_loss_acc optional: [0] += sum_r -logp_r , [1] += #non-finite rows.              */
int sbi_b200_nsf_vjp_parts(int64_t R);
int sbi_b200_nsf_vjp(const sbi_nsf_model* m, const sbi_rows* rows, const float* d_gout,
                     float g_const, float* d_logp, float* d_gpart, float* d_ginput,
                     float* d_gcond, float* d_loss_acc, void* stream);

/* x = flow^{-1}(noise | cond): sampling path.  d_noise (R, D) -> d_out (R, D);
 * d_logabsdet (R,) optional = log|det d x / d noise|. */
int sbi_b200_nsf_inverse(const sbi_nsf_model* m, const sbi_rows* rows, float* d_out,
                         float* d_logabsdet, void* stream);

/* `made` sampling (MixtureOfGaussiansMADE.sample, D sequential conditioner passes): d_input of `rows` holds
 * standard-normal draws (R, D), d_uniform (R, D) uniforms in [0, 1) that select the mixture components
 * (inverse CDF); d_out (R, D) samples in the ORIGINAL space (column 0 is the wrapper's dummy feature). */
int sbi_b200_made_sample(const sbi_nsf_model* m, const sbi_rows* rows, const float* d_uniform, float* d_out,
                         void* stream);

/* grad[p] = sum_i gpart[i][p]  (i < n_part) */
int sbi_b200_reduce_partials(const float* d_gpart, int n_part, int64_t n_params, float* d_grad,
                             void* stream);

/* Same, additionally emitting one partial of sum(grad^2) per reduction block (d_sumsq_part,
 * sbi_b200_sumsq_blocks(n_params) floats; masked-out entries excluded) so that the clip norm needs
 * no second pass over the gradient (single-GPU path; after an all-reduce the norm must be retaken). */
int sbi_b200_sumsq_blocks(int64_t n_params);
int sbi_b200_reduce_partials_norm(const float* d_gpart, int n_part, int64_t n_params, float* d_grad,
                                  const uint8_t* d_mask, float* d_sumsq_part, void* stream);

/* Epoch statistics of a validation pass (sbi/inference/trainers/base.py:1195-1225 followed by
 * assert_all_finite): d_out2[0] = -sum of the finite entries of d_logp (n), d_out2[1] = number of
 * non-finite entries.  One launch, fixed summation order. */
int sbi_b200_nll_stats(const float* d_logp, int64_t n, float* d_out2, void* stream);

/* clip_grad_norm_(max_norm) + Adam (torch defaults, no weight decay), in place.
 *   d_state: [m (n) | v (n)] ; d_step: int32 device counter (incremented here);
 *   grad_scale multiplies the gradient first (e.g. 1/world_size after an all-reduce);
 *   d_mask optional (n,) uint8: 0 = frozen entry (padding / structural zero).
 *   max_norm <= 0 disables clipping. */
int sbi_b200_adam_clip_step(float* d_params, const float* d_grad, float* d_state,
                            int32_t* d_step, const uint8_t* d_mask, int64_t n, float lr,
                            float beta1, float beta2, float eps, float max_norm,
                            float grad_scale, void* stream);
/* as above, taking the gradient's sum of squares from d_sumsq_part (n_sumsq partials) instead of
 * recomputing it */
int sbi_b200_adam_clip_step_norm(float* d_params, const float* d_grad, float* d_state,
                                 int32_t* d_step, const uint8_t* d_mask, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float max_norm,
                                 float grad_scale, const float* d_sumsq_part, int n_sumsq,
                                 void* stream);

/* ---- tensor-core bulk evaluation of the NSF (tcgen05 kind::tf32, 3xTF32 split, accumulators in
 * TMEM).  Same function as sbi_b200_nsf_logprob (NFlowsFlow.log_prob,
 * sbi/neural_nets/estimators/nflows_flow.py:77-97) for large row counts: the ResidualNet linears
 * (nflows ResidualNet; call site sbi/neural_nets/net_builders/flow.py:411-419) run on the tensor
 * cores, one row per TMEM lane, everything else (spline, LU, base density) per thread.
 *
 * The linears' weights are re-packed from the flat parameter buffer into d_tcw by
 * sbi_b200_nsf_tc_pack (call it whenever d_params changed): per coupling layer a sequence of
 * stages [hi | lo], each half a concatenation of K-major no-swizzle UMMA operand blocks
 * [K/4 slabs][N rows][4 floats] (hi = tf32-rounded weight, lo = weight - hi).
 *   d_src   (n_words,) gather map built by the host (sbi_b200/pack.py NsfLayout.tc_plan):
 *           -1 -> 0 ; s >= 0 -> hi(params[s]) ; s <= -2 -> lo(params[-2-s])
 *   d_tab   T * SBI_NSF_TC_STRIDE ints; per layer: [0] number of stages, [1] round8(n_id),
 *           then 4 ints per stage s at 4+4s: float offset into d_tcw, floats (hi+lo), N of the
 *           main block, aux (final-layer passes: first feature | n_features << 16).
 *           Stage order: initial layer, per block (Wc, W1, W2), final-layer passes of <= 2
 *           spline features (32 rows per feature).
 * Supported when H == 50, H + C <= 64, n_id <= 48, 3*KB-1 <= 32 and the shared-memory plan fits
 * (sbi_b200_nsf_tc_supported); callers use sbi_b200_nsf_logprob otherwise. */
#define SBI_NSF_TC_STRIDE 192
#define SBI_NSF_TC_MAX_STAGES 46
typedef struct {
  int32_t n_words;                 /* floats in d_tcw / entries in d_src */
  int32_t stage_cap;               /* floats of the largest stage (hi + lo) */
  const int32_t* d_src;
  const int32_t* d_tab;
  float* d_tcw;
} sbi_nsf_tc;

int sbi_b200_nsf_tc_supported(const sbi_nsf_model* m, const sbi_nsf_tc* tc);
int sbi_b200_nsf_tc_pack(const sbi_nsf_model* m, const sbi_nsf_tc* tc, void* stream);
int sbi_b200_nsf_logprob_tc(const sbi_nsf_model* m, const sbi_nsf_tc* tc, const sbi_rows* rows,
                            float* d_logp, float* d_noise, void* stream);
/* sampling direction on the same machinery: as sbi_b200_nsf_inverse (NFlowsFlow.sample,
 * sbi/neural_nets/estimators/nflows_flow.py:111-128) */
int sbi_b200_nsf_inverse_tc(const sbi_nsf_model* m, const sbi_nsf_tc* tc, const sbi_rows* rows,
                            float* d_out, float* d_logabsdet, void* stream);

/* ---- tensor-core training step (csrc/nsf_tc.cu forward sweep with activation save + csrc/nsf_vjp_tc.cu
 * backward sweep): the parameter gradients of  sum_r g_r log q(input_r | cond_r)  -- what
 * sbi_b200_nsf_vjp computes with d_ginput == d_gcond == NULL, i.e. NFlowsFlow.loss + autograd backward of a
 * training batch (sbi/neural_nets/estimators/nflows_flow.py:99-109, sbi/inference/trainers/base.py:1171-1180)
 * -- with every conditioner linear (forward, input gradient, weight gradient) on tcgen05.mma kind::tf32.
 *   tc_fwd: operand plan of the forward linears (pack.NsfLayout.tc_plan), as for sbi_b200_nsf_logprob_tc;
 *   tc_bwd: operand plan of the transposed linears (pack.NsfLayout.tc_bwd_plan), same descriptor format;
 *           both must have been packed from the current parameters (sbi_b200_nsf_tc_pack);
 *   d_gpart: (sbi_b200_nsf_vjp_tc_parts(R), n_params) partial gradients, reduce with sbi_b200_reduce_partials;
 *   d_save : caller-owned activation scratch of at least sbi_b200_nsf_vjp_tc_save_bytes(m, R) bytes
 *            (one slab per 128-row tile in flight: 3 KB per row and layer);
 *   d_logp (R) and d_loss_acc (2: sum of -log q over finite rows, number of non-finite rows) optional. */
int sbi_b200_nsf_vjp_tc_supported(const sbi_nsf_model* m, const sbi_nsf_tc* tc_fwd, const sbi_nsf_tc* tc_bwd);
int sbi_b200_nsf_vjp_tc_parts(int64_t R);
int64_t sbi_b200_nsf_vjp_tc_save_bytes(const sbi_nsf_model* m, int64_t R);
int sbi_b200_nsf_vjp_tc(const sbi_nsf_model* m, const sbi_nsf_tc* tc_fwd, const sbi_nsf_tc* tc_bwd,
                        const sbi_rows* rows, const float* d_gout, float g_const, float* d_logp,
                        float* d_gpart, float* d_loss_acc, float* d_save, int64_t save_bytes, void* stream);

/* ---- masked autoregressive flow (sbi `posterior_nn("maf")`, reference builder
 * sbi/neural_nets/net_builders/flow.py:115-209: T x [MaskedAffineAutoregressiveTransform(MADE,
 * feed-forward blocks, tanh) + RandomPermutation], z-scored input, standardised context).
 * Masked weights are stored already multiplied by their masks. */
#define SBI_MAF_LAYER_STRIDE 32
#define SBI_MAF_AFFINE 0
#define SBI_MAF_RQS 1
enum {
  SBI_M_W0 = 0,   /* [Hp][Dp] masked initial layer */
  SBI_M_B0 = 1,
  SBI_M_WC = 2,   /* [Hp][Cp] context layer */
  SBI_M_BC = 3,
  SBI_M_WF = 4,   /* [OUTp][Hp] masked final layer; rows OUTM*d .. OUTM*d+OUTM-1 parameterise feature d */
  SBI_M_BF = 5,
  SBI_M_PERM = 6, /* offset into perm_tab: perm[D] then inverse perm[D] */
  SBI_M_BLK0 = 8  /* per feed-forward block b: W at SBI_M_BLK0+2b ([Hp][Hp] masked), bias at +1 */
};
typedef struct {
  int32_t D, C, H, NB, T;
  int32_t Dp, Cp, Hp, OUTp;
  int32_t rpc0, rpc1, rpcf;        /* rows per weight chunk: initial(+context), hidden, final */
  int32_t wcap, nbuf, n_params;
  int32_t scale_softplus;          /* 1: softplus(s)+1e-3 (what sbi's maf computes), 0: sigmoid(s+2)+1e-3 */
  /* element-wise transform the MADE parameterises (`head`):
   *   SBI_MAF_AFFINE (0): OUTM = 2, row 2d = unconstrained scale, 2d+1 = shift (`maf`, flow.py:115-209)
   *   SBI_MAF_RQS    (1): OUTM = 3*KB-1 raw spline parameters per feature, linear tails (`maf_rqs`,
   *                       flow.py:212-330: MaskedPiecewiseRationalQuadraticAutoregressiveTransform) */
  int32_t head, KB, OUTM;
  float tail_bound, min_w, min_h, min_d, isq;
  float ld_zscore;
  const float* d_params;
  const int32_t* d_layer_tab;      /* T * SBI_MAF_LAYER_STRIDE */
  const int32_t* d_perm_tab;
  const float* d_stats;            /* [shift(Dp) | scale(Dp) | ctx_mean(Cp) | ctx_std(Cp)] */
} sbi_maf_model;

/* same contracts as the sbi_b200_nsf_* entry points */
int sbi_b200_maf_logprob(const sbi_maf_model* m, const sbi_rows* rows, float* d_logp,
                         float* d_noise, void* stream);
int sbi_b200_maf_vjp_parts(int64_t R);
int sbi_b200_maf_vjp(const sbi_maf_model* m, const sbi_rows* rows, const float* d_gout,
                     float g_const, float* d_logp, float* d_gpart, float* d_ginput,
                     float* d_gcond, float* d_loss_acc, void* stream);
int sbi_b200_maf_inverse(const sbi_maf_model* m, const sbi_rows* rows, float* d_out,
                         float* d_logabsdet, void* stream);

/* ---- ratio estimator: NRE `classifier_nn("resnet")` (reference builder
 * sbi/neural_nets/net_builders/classifier.py:172-235 -> nflows ResidualNet(in=Dt+Dx, out=1,
 * hidden, context=None, num_blocks, relu); wrapper sbi/neural_nets/ratio_estimators.py:132-150:
 * logit = net(cat(standardize(theta), standardize(x)))). */
enum {
  SBI_R_W0 = 0, SBI_R_B0 = 1,   /* [Hp][Dtp+Dxp], columns = [theta | pad | x | pad] */
  SBI_R_WF = 2, SBI_R_BF = 3,   /* [4][Hp] (row 0 is the logit) */
  SBI_R_BLK0 = 4                /* per block b: W1,B1,W2,B2 at SBI_R_BLK0 + 4b */
};
typedef struct {
  int32_t Dt, Dx, H, NB;
  int32_t Dtp, Dxp, Hp;
  int32_t rpc0, rpc1;
  int32_t wcap, nbuf, n_params;
  const float* d_params;
  const int32_t* d_tab;           /* SBI_R_* offsets */
  const float* d_stats;           /* [theta_mean(Dtp) | theta_std(Dtp) | x_mean(Dxp) | x_std(Dxp)] */
} sbi_ratio_model;

/* rows of (theta, x) pairs: pair r = (theta[ti[r]], x[xi[r]]); NULL index = identity;
 * x_shared = 1: every pair uses x row 0 (potential at a fixed observation). */
typedef struct {
  const float* d_theta;
  const float* d_x;
  const int64_t* d_theta_index;
  const int64_t* d_x_index;
  int64_t R;
  int32_t x_shared;
} sbi_pairs;

/* unnormalised log-ratio logits (R,) -- RatioEstimator.forward, ratio_estimators.py:132-154 */
int sbi_b200_ratio_forward(const sbi_ratio_model* m, const sbi_pairs* pairs, float* d_logits,
                           void* stream);
/* VJP of sum_r g_r logit_r: d_gpart (n_part, n_params) per-CTA partial parameter gradients,
 * d_gtheta (R, Dt) optional gradient wrt the theta of each pair; d_logits optional. */
int sbi_b200_ratio_vjp_parts(int64_t R);
int sbi_b200_ratio_vjp(const sbi_ratio_model* m, const sbi_pairs* pairs, const float* d_gout,
                       float* d_logits, float* d_gpart, float* d_gtheta, void* stream);

/* tensor-core bulk evaluation of the classifier (same operand format and `sbi_nsf_tc` descriptor as
 * the NSF path: d_tab holds ONE stage list: initial layer, per block (W1, W2), final layer as an
 * N = 16 block whose row 0 is the weight vector; [1] = round8(Dt + Dx)).  Supported when H == 50 and
 * Dt + Dx <= 56.  Gather map: sbi_b200/pack.py RatioLayout.tc_plan. */
int sbi_b200_ratio_tc_supported(const sbi_ratio_model* m, const sbi_nsf_tc* tc);
int sbi_b200_ratio_tc_pack(const sbi_ratio_model* m, const sbi_nsf_tc* tc, void* stream);
int sbi_b200_ratio_forward_tc(const sbi_ratio_model* m, const sbi_nsf_tc* tc, const sbi_pairs* pairs,
                              float* d_logits, void* stream);

/* ---- lock-step vectorized slice sampler (state machine of
 * sbi/samplers/mcmc/slice_numpy.py:412-587 `SliceSamplerVectorized.run`): one thread per chain,
 * chain state resident in HBM, one launch per lock-step between two potential evaluations.
 * Coordinate-wise slice sampling with stepping-out (bracket width tuned as the running mean of the
 * bracket sizes during the first `tuning` sweeps) and shrinkage; random dimension order per sweep. */
enum { SBI_SLICE_BEGIN = 0, SBI_SLICE_LOWER = 1, SBI_SLICE_UPPER = 2, SBI_SLICE_SAMPLE = 3, SBI_SLICE_DONE = 4 };
typedef struct {
  int32_t C, D;                 /* chains, dimensions */
  int32_t num_samples, tuning;  /* sweeps to record per chain, tuning sweeps before recording */
  double init_width, max_width;
  uint64_t seed;
  double* d_x;                  /* (C, D) current position (in/out) */
  double* d_width;              /* (C, D) */
  int32_t* d_order;             /* (C, D) */
  int32_t* d_istate;            /* (C, 4): state, i, t, - */
  double* d_fstate;             /* (C, 8): cxi, wi, lx, ux, xi, logu, -, - */
  void* d_rng;                  /* (C, 64 bytes) Philox state */
  double* d_samples;            /* (C, num_samples, D) */
} sbi_slice_chains;

/* initialise chain state from d_x; writes the first parameters to evaluate into d_params (C, D) f32 */
int sbi_b200_slice_init(const sbi_slice_chains* s, float* d_params, void* stream);
/* one lock-step: consume d_logp (C,) evaluated at d_params, advance every chain, write the next
 * d_params; d_n_done[0] = number of chains in state DONE after this step */
int sbi_b200_slice_step(const sbi_slice_chains* s, const float* d_logp, float* d_params,
                        int32_t* d_n_done, void* stream);

/* ---- flow matching (FMPE): VectorFieldMLP behind FlowMatchingEstimator
 * (net: sbi/neural_nets/net_builders/vector_field_nets.py:610-719, sinusoidal time embedding
 * :367-421; estimator: sbi/neural_nets/estimators/flowmatching_estimator.py:205-347). */
#define SBI_FM_MAX_LAYERS 12
enum {
  SBI_F_WI = 0, SBI_F_BI = 1,    /* input_layer        [Hp][Dp] */
  SBI_F_WC = 2, SBI_F_BC = 3,    /* condition_layer    [Hp][Cp] */
  SBI_F_WM = 4, SBI_F_BM = 5,    /* input_merge_layer  [Hp][2Hp], columns = [input emb | cond emb] */
  SBI_F_WT = 6, SBI_F_BT = 7,    /* time_linear_layer  [Hp][TEp] */
  SBI_F_WO = 8, SBI_F_BO = 9,    /* output_layer       [Dp][Hp] */
  SBI_F_LAYER0 = 12              /* per hidden layer i: W, B, LN gamma, LN beta at SBI_F_LAYER0 + 4i */
};
typedef struct {
  int32_t D, C, H, NL, TE;          /* theta dim, (embedded) condition dim, hidden, layers, time-emb dim */
  int32_t Dp, Cp, Hp, TEp;
  int32_t rpc_i, rpc_c, rpc_m, rpc_t, rpc_h, rpc_o;   /* rows per weight chunk of each matrix */
  int32_t wcap, nbuf, n_params;
  float noise_scale;                /* sigma_min = 1e-3 */
  float ln_eps;
  int32_t raw;                      /* 1: the bare network (score estimators): d_input is already the network
                                     * input, d_time the value fed to the time embedding, outputs are the raw
                                     * network outputs -- no flow-matching noising / standardisation / rescaling */
  int32_t pad_;
  const float* d_params;
  const int32_t* d_tab;
  const float* d_stats;             /* [mean_0(Dp) | std_0(Dp) | ctx_mean(Cp) | ctx_std(Cp) | div_term(TEp/2)] */
} sbi_fm_model;

/* v(theta_t, t; x) in ORIGINAL space (FlowMatchingEstimator.forward :205-268): d_theta (R,D),
 * d_cond (R,C) or (1,C) when cond_shared, d_time (R,) or (1,) when time_shared -> d_v (R,D). */
int sbi_b200_fm_forward(const sbi_fm_model* m, const sbi_rows* rows, const float* d_time,
                        int32_t time_shared, float* d_v, void* stream);
/* the same velocity field and its exact divergence sum_i dv_i/dtheta_i (d_div (R,); d_v optional): the
 * right-hand side of the augmented neural ODE behind `VectorFieldPosterior.log_prob`
 * (sbi/samplers/ode_solvers/zuko_ode.py:80-124 -> zuko FreeFormJacobianTransform(exact=True);
 * sbi/inference/potentials/vector_field_potential.py:145-212).  With m->raw the kernel writes the DIAGONAL of the bare
 * network's input Jacobian instead, d_div (R, D): the caller (score estimators) weights it per dimension. */
int sbi_b200_fm_forward_div(const sbi_fm_model* m, const sbi_rows* rows, const float* d_time,
                            int32_t time_shared, float* d_v, float* d_div, void* stream);
/* flow-matching loss of a batch and its parameter gradient (FlowMatchingEstimator.loss :270-347
 * + backward): rows = (theta_0, x) pairs, d_time (R,) in [0,1], d_eps (R,D) ~ N(0,I).
 * d_loss (R,) optional; d_gpart (n_part, n_params) partial gradients of sum_r g_r * loss_r. */
int sbi_b200_fm_vjp_parts(int64_t R);
int sbi_b200_fm_loss_vjp(const sbi_fm_model* m, const sbi_rows* rows, const float* d_time,
                         const float* d_eps, const float* d_gout, float g_const, float* d_loss,
                         float* d_gpart, float* d_loss_acc, void* stream);

/* Introspection, no device work: the weight-pipeline plan (ring depth, chunk rows, shared memory) that a launch of
 * kernel 0 (sbi_b200_fm_forward), 1 (sbi_b200_fm_loss_vjp / sbi_b200_fm_net_vjp) or 2 (sbi_b200_fm_forward_div)
 * uses for this model: out10 = [nbuf, wcap, rpc_i, rpc_c, rpc_m, rpc_t, rpc_h, rpc_o, dynamic smem bytes, output
 * rows per thread].  The launches re-chunk the caller's plan to fill the 227 KB of shared memory. */
int sbi_b200_fm_plan(const sbi_fm_model* m, int32_t kernel, int32_t* out10);

/* Parameter gradient of the bare network for a given upstream gradient d_dout (R, D) of its outputs (m->raw must
 * be 1): the backward of `ConditionalScoreEstimator.forward` (sbi/neural_nets/estimators/score_estimator.py:149-215)
 * through the VectorFieldMLP; everything around the network (time-dependent z-scoring, the Gaussian skip term, the
 * denoising-score-matching loss with its control variate, :230-316) is element-wise host code.
 * d_gpart (sbi_b200_fm_vjp_parts(R), n_params) receives per-CTA partial gradients. */
int sbi_b200_fm_net_vjp(const sbi_fm_model* m, const sbi_rows* rows, const float* d_time, const float* d_dout,
                        float* d_gpart, void* stream);

/* ---- adaptive Dormand-Prince 5(4) with the step control on the device (csrc/ode.cu), replacing the
 * host-side loop of the solver the reference delegates to (zuko.utils.odeint; call sites
 * sbi/samplers/ode_solvers/zuko_ode.py:80-124, sbi/inference/posteriors/vector_field_posterior.py:436-505).
 * The state is a flat fp32 vector of n entries; d_k holds the 7 stage derivatives (7 x n).  One step =
 * k_0 given (first-same-as-last), for i = 1..6: sbi_b200_ode_stage(i) then the caller's right-hand side at
 * time d_ctrl->t_stage into k_i; then sbi_b200_ode_error_commit.  Before the first step: stage 0 (copies y,
 * sets t_stage = t) and the right-hand side into k_0.  The host only polls d_ctrl->done. */
typedef struct {
  float t, h, t1, dir;      /* clock, signed step, end time, +1 / -1 */
  float atol, rtol;
  float t_stage;            /* time of the stage just prepared (input of the right-hand side) */
  float en;                 /* error norm of the last step */
  int32_t nfe, nsteps, naccept;
  int32_t done;             /* 0 running, 1 reached t1, 2 max_steps exhausted */
  int32_t max_steps;
  int32_t pad_[3];
} sbi_ode_ctrl;
int sbi_b200_ode_red_size(int64_t n);        /* floats of scratch the error reduction needs */
int sbi_b200_ode_stage(const float* d_y, const float* d_k, float* d_yi, int64_t n, int32_t stage,
                       sbi_ode_ctrl* d_ctrl, void* stream);
int sbi_b200_ode_error_commit(float* d_y, float* d_k, float* d_y5, float* d_red, int64_t n,
                              sbi_ode_ctrl* d_ctrl, void* stream);

/* One Euler-Maruyama step of the reverse SDE of the flow-matching estimator (csrc/ode.cu; reference
 * sbi/samplers/score/predictors.py:112-120 on flowmatching_estimator.py:374-469): d_theta (n = R*D) is
 * updated in place from the velocity d_v at time ts[i-1] and the normal draw d_z; d_ctrl = [current time,
 * step index i] (floats) is advanced to ts[i], i+1, so that a captured step can be replayed. */
int sbi_b200_sde_em_step(float* d_theta, const float* d_v, const float* d_z, int64_t n, const float* d_ts,
                         float* d_ctrl, float eta, float noise_scale, float t_eff, void* stream);

/* ---- rejection sampling: accept / reject + order-preserving compaction of one batch of proposals
 * (csrc/compact.cu; reference sbi/samplers/rejection/rejection.py:170-200, `keep = exp(potential - scaled
 * proposal log-prob) > u; candidates[keep]`).  Accepted rows are appended, in proposal order, at
 * d_out[*d_count ...] (rows past `cap` are dropped but counted), their global proposal indices
 * (index_base + position) at d_out_idx (optional); *d_count is advanced on the device.
 * d_scratch: sbi_b200_reject_scratch_ints(n) int32. */
int64_t sbi_b200_reject_scratch_ints(int64_t n);
int sbi_b200_reject_compact(const float* d_cand, int32_t D, const float* d_log_target, const float* d_log_scaled,
                            const float* d_u, int64_t n, int64_t index_base, float* d_out, int64_t* d_out_idx,
                            int64_t cap, int32_t* d_count, int32_t* d_scratch, void* stream);

/* ---- multi-GPU: gradient sum over NVLink peer memory (csrc/peer.cu), replacing the NCCL all-reduce +
 * norm pass of the data-parallel step (reference semantics: clip_grad_norm_ + Adam on the summed
 * gradient, sbi/inference/trainers/base.py:1181-1187).  Each rank allocates a symmetric buffer
 * (`peer_alloc`), exports its IPC handle (64 bytes) to the other ranks of the node, imports theirs, and
 * then calls `peer_sum` once per step with the table of all ranks' buffers (own buffer at [rank]):
 * d_grad_out = sum over ranks (fixed rank order) of d_grad_local, plus sbi_b200_peer_blocks(n) partials
 * of sum(g^2) for sbi_b200_adam_clip_step_norm.  The step number that tags the flags is read from the
 * exchange's own device counter inside the symmetric buffer (d_step == NULL; advanced by the kernel, never
 * rewound) or from d_step[0] (a caller-owned counter, e.g. the optimizer's), so the launch can sit in a
 * CUDA graph. */
int64_t sbi_b200_peer_bytes(int64_t n_params);
int sbi_b200_peer_blocks(int64_t n_params);
void* sbi_b200_peer_alloc(int64_t n_params);
int sbi_b200_peer_free(void* p);
int sbi_b200_peer_export(void* p, void* handle64);
void* sbi_b200_peer_import(const void* handle64);
int sbi_b200_peer_close(void* p);
int sbi_b200_peer_sum(const float* d_grad_local, void* const* h_peer_ptrs, int world, int rank,
                      int64_t n_params, float* d_grad_out, const uint8_t* d_mask, float* d_sumsq_part,
                      const int32_t* d_step, void* stream);
int sbi_b200_peer_error(const void* p, int64_t n_params);

/* ---- host-buffer entry points (the end-to-end path a CPU caller binds) ------------------
 * Device staging / optimizer buffers are owned by the caller and passed in a workspace;
 * h_* buffers should be pinned for full PCIe bandwidth.  These calls copy host->device,
 * run the kernels, copy the result device->host and block until it has landed. */
typedef struct {
  float* d_input;        /* (cap_rows, D) staging */
  float* d_cond;         /* (cap_rows, C) staging */
  float* d_logp;         /* (cap_rows) */
  float* d_gpart;        /* (sbi_b200_nsf_vjp_parts(cap_rows), n_params), zero-initialised once */
  float* d_grad;         /* (n_params) */
  float* d_state;        /* (2*n_params) Adam m | v */
  int32_t* d_step;       /* (2) */
  const uint8_t* d_mask; /* (n_params) or NULL */
  float* d_loss_acc;     /* (2) */
  int64_t cap_rows;
  float* d_sumsq;        /* (sbi_b200_sumsq_blocks(n_params)) scratch for the clip norm, or NULL */
  /* optional: run the step's forward+backward on the tensor cores (sbi_b200_nsf_vjp_tc); all NULL / 0
   * selects the SIMT kernel.  tc_pack covers both operand plans (one sbi_b200_nsf_tc_pack launch per
   * step re-packs them from the just-updated parameters); d_gpart must then hold
   * sbi_b200_nsf_vjp_tc_parts(B) slabs and d_save sbi_b200_nsf_vjp_tc_save_bytes(m, B) bytes. */
  const sbi_nsf_tc* tc_pack;
  const sbi_nsf_tc* tc_fwd;
  const sbi_nsf_tc* tc_bwd;
  float* d_save;
  int64_t save_bytes;
} sbi_train_ws;

/* One optimisation step on a host batch (replaces one iteration of
 * sbi/inference/trainers/base.py:1171-1187 incl. the batch `.to(device)` of
 * npe_base.py:722-726): loss = mean_r -log q(theta_r | x_r); h_loss_out[0] = sum_r -log q,
 * h_loss_out[1] = number of non-finite rows. */
int sbi_b200_nsf_train_step_host(const sbi_nsf_model* m, const sbi_train_ws* ws,
                                 const float* h_input, const float* h_cond, int64_t B, float lr,
                                 float beta1, float beta2, float eps, float max_norm,
                                 float* h_loss_out, void* stream);

/* Pipelined variant: enqueues step i (H2D of its pinned host batch, kernels, D2H of its loss) and
 * returns after step i-1 has completed, handing back step i-1's result in h_loss_prev[2] (NaN on
 * the first call).  Every step still carries its own H2D + D2H; the host just prepares batch i+1
 * while the device runs step i.  The caller alternates between two pinned host batch buffers
 * (buffer i%2 may be rewritten once call i+1 has returned).  `pipe` from sbi_b200_pipe_create. */
void* sbi_b200_pipe_create(void);
void sbi_b200_pipe_destroy(void* pipe);
int sbi_b200_nsf_train_step_host_async(const sbi_nsf_model* m, const sbi_train_ws* ws, void* pipe,
                                       const float* h_input, const float* h_cond, int64_t B, float lr,
                                       float beta1, float beta2, float eps, float max_norm,
                                       float* h_loss_prev, void* stream);
/* wait for the last enqueued step and return its result in h_loss_last[2] */
int sbi_b200_pipe_drain(void* pipe, float* h_loss_last);

/* Data-parallel pipelined host step (one process per GPU): as sbi_b200_nsf_train_step_host_async, with the
 * gradient sum over NVLink peer memory (sbi_b200_peer_sum, exchange-owned step counter) between the
 * partial-gradient reduction and clip+Adam; rows of all ranks form one global batch of B * world rows
 * (reference semantics: one optimizer step on the mean loss of the global batch,
 * sbi/inference/trainers/base.py:1171-1187).  ws->d_sumsq must hold sbi_b200_peer_blocks(n_params) floats. */
typedef struct {
  void* const* h_peer_ptrs;   /* (world) symmetric buffers, own buffer at [rank] (host array) */
  int world, rank;
  float* d_grad_local;        /* (n_params) scratch for this rank's reduced gradient */
} sbi_peer_ctx;
int sbi_b200_nsf_train_step_host_async_dp(const sbi_nsf_model* m, const sbi_train_ws* ws, void* pipe,
                                          const sbi_peer_ctx* peer, const float* h_input, const float* h_cond,
                                          int64_t B, float lr, float beta1, float beta2, float eps,
                                          float max_norm, float* h_loss_prev, void* stream);

/* log q(input_r | cond) for R host rows (cond: (R,C), or (1,C) when cond_shared). */
int sbi_b200_nsf_logprob_host(const sbi_nsf_model* m, const sbi_train_ws* ws,
                              const float* h_input, const float* h_cond, int64_t R,
                              int cond_shared, float* h_logp, void* stream);

/* same through the tensor-core kernel (sbi_b200_nsf_logprob_tc), in chunks whose host<->device
 * copies overlap the kernels of their neighbours; re-packs tc->d_tcw first. */
int sbi_b200_nsf_logprob_host_tc(const sbi_nsf_model* m, const sbi_nsf_tc* tc, const sbi_train_ws* ws,
                                 const float* h_input, const float* h_cond, int64_t R,
                                 int cond_shared, float* h_logp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SBI_B200_H */
