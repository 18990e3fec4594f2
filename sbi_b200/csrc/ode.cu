// Adaptive Dormand-Prince 5(4) with the step control on the device: the stage combinations, the error
// norm, the accept / reject decision and the next step size never leave the GPU, so an ODE solve of
// the flow-matching posterior (sampling, and the neural-ODE log-probability with its exact trace) is
// a fixed launch sequence per step -- captured once as a CUDA graph and replayed -- with one host
// read of a `done` flag every few steps instead of one `.item()` per step.
//
// Restates the solver the reference delegates to (zuko.utils.odeint, third party, pinned zuko==1.6.0
// in /root/reference/uv.lock; call sites /root/reference/sbi/samplers/ode_solvers/zuko_ode.py:80-124,
// /root/reference/sbi/inference/posteriors/vector_field_posterior.py:436-505): one step size for the
// whole batch, error norm = RMS over ALL state entries of err / (atol + rtol * max(|y|, |y_new|)),
// accept iff <= 1, step factor 0.9 * err^(-1/5) clamped to [0.2, 5], first-same-as-last stage reuse.
// The state is one flat fp32 vector (for log_prob: [theta (R, D) | log|det| (R)]).
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/sbi_b200.h"
#include "device.cuh"

namespace sbi {
namespace ode {

// Butcher tableau (Dormand & Prince 1980)
__constant__ float cA[7][6] = {
    {0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {1.f / 5, 0.f, 0.f, 0.f, 0.f, 0.f},
    {3.f / 40, 9.f / 40, 0.f, 0.f, 0.f, 0.f},
    {44.f / 45, -56.f / 15, 32.f / 9, 0.f, 0.f, 0.f},
    {19372.f / 6561, -25360.f / 2187, 64448.f / 6561, -212.f / 729, 0.f, 0.f},
    {9017.f / 3168, -355.f / 33, 46732.f / 5247, 49.f / 176, -5103.f / 18656, 0.f},
    {35.f / 384, 0.f, 500.f / 1113, 125.f / 192, -2187.f / 6784, 11.f / 84}};
__constant__ float cC[7] = {0.f, 1.f / 5, 3.f / 10, 4.f / 5, 8.f / 9, 1.f, 1.f};
__constant__ float cB5[7] = {35.f / 384, 0.f, 500.f / 1113, 125.f / 192, -2187.f / 6784, 11.f / 84, 0.f};
__constant__ float cE[7] = {35.f / 384 - 5179.f / 57600,    0.f,
                            500.f / 1113 - 7571.f / 16695,  125.f / 192 - 393.f / 640,
                            -2187.f / 6784 + 92097.f / 339200, 11.f / 84 - 187.f / 2100,
                            -1.f / 40};

constexpr int kBlock = 256;

// y_i = y + h * sum_j a_ij k_j ; t_stage = t + c_i h   (stage 0: y_i = y, t_stage = t)
__global__ void stage_kernel(const float* __restrict__ y, const float* __restrict__ k, float* __restrict__ yi,
                             int64_t n, int stage, sbi_ode_ctrl* __restrict__ c) {
  const float h = c->h;
  if (blockIdx.x == 0 && threadIdx.x == 0) c->t_stage = c->t + cC[stage] * h;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < stage; ++j) a = fmaf(cA[stage][j], k[(int64_t)j * n + i], a);
    yi[i] = fmaf(h, a, y[i]);
  }
}

// y5 = y + h sum b5_j k_j ; err = h sum (b5_j - b4_j) k_j ; red[block] = sum (err / tol)^2
__global__ void error_kernel(const float* __restrict__ y, const float* __restrict__ k, float* __restrict__ y5,
                             float* __restrict__ red, int64_t n, const sbi_ode_ctrl* __restrict__ c) {
  __shared__ float sh[kBlock];
  const float h = c->h, atol = c->atol, rtol = c->rtol;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s5 = 0.f, se = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const float kj = k[(int64_t)j * n + i];
      s5 = fmaf(cB5[j], kj, s5);
      se = fmaf(cE[j], kj, se);
    }
    const float y0 = y[i];
    const float yn = fmaf(h, s5, y0);
    y5[i] = yn;
    const float q = (h * se) / (atol + rtol * fmaxf(fabsf(y0), fabsf(yn)));
    acc = fmaf(q, q, acc);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kBlock / 2; o > 0; o >>= 1) {      // fixed tree: deterministic
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) red[blockIdx.x] = sh[0];
}

// accept / reject (every block derives the same decision from the same partials), commit the state and
// the first-same-as-last stage, then block 0 advances the clock and picks the next step size.
__global__ void commit_kernel(float* __restrict__ y, const float* __restrict__ y5, float* __restrict__ k,
                              const float* __restrict__ red, int nred, int64_t n, sbi_ode_ctrl* __restrict__ c) {
  __shared__ float s_en;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < nred; ++b) s += red[b];
    float en = sqrtf(s / (float)n);
    if (!(en == en)) en = INFINITY;                 // NaN -> reject and shrink
    s_en = en;
  }
  __syncthreads();
  const float en = s_en;
  const bool accept = en <= 1.f;
  if (accept) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      y[i] = y5[i];
      k[i] = k[6 * n + i];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && !c->done) {
    float t = c->t, h = c->h;
    c->en = en;
    c->nsteps += 1;
    c->nfe += 6;
    if (accept) { t += h; c->naccept += 1; }
    const float fac = 0.9f * powf(1.f / fmaxf(en, 1e-10f), 0.2f);
    h *= fminf(5.f, fmaxf(0.2f, fac));
    const float dir = c->dir;
    if ((c->t1 - t) * dir <= 1e-12f) {
      c->done = 1;
      t = c->t1;
      h = 0.f;
    } else if ((t + h - c->t1) * dir > 0.f) {
      h = c->t1 - t;
    }
    if (c->nsteps >= c->max_steps) { c->done = 2; h = 0.f; }
    c->t = t;
    c->h = h;
  }
}

// ---- reverse SDE, Euler-Maruyama predictor (no corrector) ---------------------------------------------------
// One step of /root/reference/sbi/samplers/score/predictors.py:112-120 (EulerMaruyama.predict) on the
// flow-matching estimator's SDE view (/root/reference/sbi/neural_nets/estimators/flowmatching_estimator.py:
// 374-469: score, drift_fn, diffusion_fn), driven by Diffuser.run (samplers/score/diffuser.py:124-180):
//   f = -theta / max(1 - t1, 1 - t_eff) ; g = sqrt(2 (t1 + s) / max(1 - t1, 1 - t_eff))
//   score = (-(1 - t1) v - theta) / (t1 + s) ; theta <- theta - (f - (1 + eta^2)/2 g^2 score) dt + eta g z sqrt(dt)
// with t1 = ts[i-1], dt = t1 - ts[i], v the velocity at (theta, t1) (fm_forward reads t1 from ctrl[0]).
// ctrl = [t_cur, step index as float]; sde_advance moves it to the next grid point after the update.
__global__ void sde_em_step_kernel(float* __restrict__ theta, const float* __restrict__ v, const float* __restrict__ z,
                                   int64_t n, const float* __restrict__ ts, const float* __restrict__ ctrl, float eta,
                                   float noise_scale, float t_eff) {
  const int i = (int)ctrl[1];
  const float t1 = ts[i - 1], t0 = ts[i];
  const float dt = t1 - t0;
  const float omt = fmaxf(1.f - t1, 1.f - t_eff);
  const float g = sqrtf(2.f * (t1 + noise_scale) / omt);
  const float c = (1.f + eta * eta) / 2.f * g * g;
  const float sq = sqrtf(dt);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float th = theta[e];
    const float f = -th / omt;
    const float score = (-(1.f - t1) * v[e] - th) / (t1 + noise_scale);
    const float fb = f - c * score;
    theta[e] = th - fb * dt + (eta * g) * z[e] * sq;
  }
}
__global__ void sde_advance_kernel(const float* __restrict__ ts, float* __restrict__ ctrl) {
  const int i = (int)ctrl[1];
  ctrl[0] = ts[i];
  ctrl[1] = (float)(i + 1);
}

static int grid_for(int64_t n) {
  const int64_t b = (n + kBlock - 1) / kBlock;
  const int64_t cap = (int64_t)dev_num_sms() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace ode
}  // namespace sbi

using namespace sbi;

extern "C" int sbi_b200_ode_red_size(int64_t n) { return n < 1 ? 1 : ode::grid_for(n); }

extern "C" int sbi_b200_ode_stage(const float* d_y, const float* d_k, float* d_yi, int64_t n, int32_t stage,
                                  sbi_ode_ctrl* d_ctrl, void* stream) {
  sbi::DeviceGuard dev_guard_(d_y);
  if (!d_y || !d_k || !d_yi || !d_ctrl || n < 1 || stage < 0 || stage > 6) return SBI_EINVAL;
  ode::stage_kernel<<<ode::grid_for(n), ode::kBlock, 0, (cudaStream_t)stream>>>(d_y, d_k, d_yi, n, stage, d_ctrl);
  return (int)cudaGetLastError();
}

extern "C" int sbi_b200_ode_error_commit(float* d_y, float* d_k, float* d_y5, float* d_red, int64_t n,
                                         sbi_ode_ctrl* d_ctrl, void* stream) {
  sbi::DeviceGuard dev_guard_(d_y);
  if (!d_y || !d_k || !d_y5 || !d_red || !d_ctrl || n < 1) return SBI_EINVAL;
  const int g = ode::grid_for(n);
  ode::error_kernel<<<g, ode::kBlock, 0, (cudaStream_t)stream>>>(d_y, d_k, d_y5, d_red, n, d_ctrl);
  ode::commit_kernel<<<g, ode::kBlock, 0, (cudaStream_t)stream>>>(d_y, d_y5, d_k, d_red, g, n, d_ctrl);
  return (int)cudaGetLastError();
}

extern "C" int sbi_b200_sde_em_step(float* d_theta, const float* d_v, const float* d_z, int64_t n, const float* d_ts,
                                    float* d_ctrl, float eta, float noise_scale, float t_eff, void* stream) {
  sbi::DeviceGuard dev_guard_(d_theta);
  if (!d_theta || !d_v || !d_z || !d_ts || !d_ctrl || n < 1) return SBI_EINVAL;
  ode::sde_em_step_kernel<<<ode::grid_for(n), ode::kBlock, 0, (cudaStream_t)stream>>>(d_theta, d_v, d_z, n, d_ts, d_ctrl,
                                                                                     eta, noise_scale, t_eff);
  ode::sde_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_ts, d_ctrl);
  return (int)cudaGetLastError();
}
