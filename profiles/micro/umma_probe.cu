// tcgen05 probe for the layouts the tensor-core NSF kernel relies on (sm_100a).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
//
// Checks, each against a host computation:
//   1. TMEM round trip: tcgen05.st.32x32b.xN of per-thread registers, tcgen05.ld back.
//   2. kind::tf32 MMA with A (128 x K) taken from TMEM (one row per lane, one element per
//      column), B (N x K, K-major) from shared memory in the no-swizzle canonical layout
//      [K/4 slabs][N rows][4 floats]: core matrix = 8 rows x 16 B contiguous,
//      SBO (8-row group stride) = 128 B, LBO (K-adjacent core matrix stride) = N*16 B.
//      Run with both (LBO,SBO) assignments to pin which descriptor field is which.
//   3. 3xTF32 split accuracy (hi = truncated fp32, lo = x - hi) on random fp32 data.
//   4. MMA issue + commit + wait round-trip latency and 3xTF32 layer time.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T, one K=8 step
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d, uint32_t a, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  // c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10, K-major A and B,
  // N>>3 @17, M>>4 @24   (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ inline uint64_t make_bdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base offset 0, lbo mode 0, SWIZZLE_NONE
}

constexpr int M = 128;

// mode 0: single product A*B^T     (values exactly representable in tf32 -> exact)
// mode 1: 3xTF32 of fp32 data
// out[M][N]; rt[M][32] TMEM round-trip check; cyc[0] = clocks of `reps` layers
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ Bc,
                                             float* __restrict__ out, float* __restrict__ rt,
                                             long long* __restrict__ cyc, int N, int K, int mode,
                                             int swap, int reps) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* bh = reinterpret_cast<float*>(smem);           // [K/4][N][4]
  float* bl = bh + N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < N * K; i += 128) {
    float w = Bc[i];
    float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    bh[i] = (mode == 0) ? w : hi;
    bl[i] = w - hi;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // generic-proxy writes of B must be visible to the tensor core's async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tbase_s)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tbase_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t tA_hi = tbase + 0, tA_lo = tbase + 64, tD = tbase + 128;   // columns

  // --- 1. round trip ---
  {
    float v[8], r[8];
    for (int c = 0; c < 32; c += 8) {
      for (int i = 0; i < 8; ++i) v[i] = (float)(tid * 100 + c + i);
      tc_st8(tD + lane_base + c, v);
    }
    tc_wait_st();
    for (int c = 0; c < 32; c += 8) {
      tc_ld8(tD + lane_base + c, r);
      for (int i = 0; i < 8; ++i) rt[tid * 32 + c + i] = r[i];
    }
  }

  // --- 2/3. A -> TMEM (row = lane) ---
  for (int c = 0; c < K; c += 8) {
    float hi[8], lo[8];
    for (int i = 0; i < 8; ++i) {
      float a = A[tid * K + c + i];
      float h = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
      hi[i] = (mode == 0) ? a : h;
      lo[i] = a - h;
    }
    tc_st8(tA_hi + lane_base + c, hi);
    tc_st8(tA_lo + lane_base + c, lo);
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();

  const uint32_t idesc = make_idesc(M, N);
  const uint32_t slab = (uint32_t)N * 16u;       // bytes of one 4-wide K slab
  const uint32_t lbo = swap ? 128u : slab, sbo = swap ? slab : 128u;
  long long t0 = 0, t1 = 0;
  uint32_t parity = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (rep == 1) t0 = clock64();
    if (tid == 0) {
      tc_fence_after();
      uint32_t acc = 0;
      for (int k = 0; k < K; k += 8) {
        const uint32_t boff = (uint32_t)(k / 4) * slab;
        tc_mma_tf32_ts(tD, tA_hi + k, make_bdesc(smem_u32(bh) + boff, lbo, sbo), idesc, acc);
        acc = 1;
        if (mode >= 1) {
          tc_mma_tf32_ts(tD, tA_lo + k, make_bdesc(smem_u32(bh) + boff, lbo, sbo), idesc, 1);
          tc_mma_tf32_ts(tD, tA_hi + k, make_bdesc(smem_u32(bl) + boff, lbo, sbo), idesc, 1);
        }
      }
      if (mode != 2 || rep == reps - 1) tc_commit(&bar);
    }
    if (mode != 2 || rep == reps - 1) {
      mbar_wait(&bar, parity);
      parity ^= 1u;
      tc_fence_after();
    }
  }
  t1 = clock64();
  if (tid == 0) cyc[0] = (reps > 1) ? (t1 - t0) / (reps - 1) : 0;

  for (int c = 0; c < N; c += 8) {
    float r[8];
    tc_ld8(tD + lane_base + c, r);
    for (int i = 0; i < 8; ++i) out[tid * N + c + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512)
                 : "memory");
}


// ---- cadence of the tensor pipe for the shapes the NSF kernel uses ------------------------------
// var 0: one accumulation chain (same D)          var 1: two chains alternating (D0 / D1)
// var 2: A from shared memory (SS), one chain     var 3: four chains
__device__ __forceinline__ void tc_mma_tf32_ss(uint32_t d, uint64_t adesc, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__global__ void __launch_bounds__(128) cadence(long long* __restrict__ cyc, int N, int var, int nmma) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* sf = reinterpret_cast<float*>(smem);
  for (int i = tid; i < (256 * 64 + 128 * 64); i += 128) sf[i] = 0.001f * (float)(i % 97);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tbase_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tbase_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = 0.01f * (float)(tid + i);
    for (int c = 0; c < 64; c += 8) tc_st8(tbase + lane_base + c, v);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  const uint32_t idesc = make_idesc(128, N);
  const uint32_t slab = (uint32_t)N * 16u;
  const uint32_t bsm = smem_u32(smem);
  const uint32_t asm_ = bsm + 256 * 64 * 4;       // A tile in smem for the SS variant: [K/4][128][4]
  long long t0 = 0, t1 = 0;
  if (var >= 4) {
    // tight issue: descriptors precomputed, K-steps unrolled; var 5: whole warp converged + elect
    uint64_t bd[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bd[k] = make_bdesc(bsm + (uint32_t)(2 * k) * slab, slab, 128u);
    const uint32_t d = tbase + 64, a = tbase;
    if (var == 4) {
      if (tid == 0) {
        tc_fence_after();
        t0 = clock64();
        for (int i = 0; i < nmma; i += 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k) tc_mma_tf32_ts(d, a + 8 * k, bd[k], idesc, 1u);
        }
        tc_commit(&bar);
      }
    } else if (warp == 0) {
      tc_fence_after();
      uint32_t leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      t0 = clock64();
      for (int i = 0; i < nmma; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (leader) tc_mma_tf32_ts(d, a + 8 * k, bd[k], idesc, 1u);
      }
      if (leader) tc_commit(&bar);
    }
  } else
  if (tid == 0) {
    tc_fence_after();
    t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const int k = (i % 8);
      const uint64_t bd = make_bdesc(bsm + (uint32_t)(2 * k) * slab, slab, 128u);
      uint32_t d = tbase + 64;
      if (var == 1) d = tbase + 64 + (i & 1) * 192;
      if (var == 3) d = tbase + 64 + (i & 3) * 96;
      if (var == 2) {
        const uint64_t ad = make_bdesc(asm_ + (uint32_t)(2 * k) * 2048u, 2048u, 128u);
        tc_mma_tf32_ss(d, ad, bd, idesc, i >= 4);
      } else {
        tc_mma_tf32_ts(d, tbase + 8 * k, bd, idesc, i >= 4);
      }
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  t1 = clock64();
  if (tid == 0) cyc[0] = (t1 - t0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512)
                 : "memory");
}


// ---- operands for the weight-gradient GEMM dW = dY^T X (K = rows): both operands from shared
// memory, MN-major (the row index is K).  Canonical no-swizzle MN-major layout tried here
// (cute/atom/mma_traits_sm100.hpp, Major-MN INTERLEAVE: ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), T = 4):
//   core matrix = 8 K-rows x 4 MN-elements, 128 B contiguous (K-row j0 at +16*j0 B);
//   SBO = byte distance between consecutive groups of 4 along MN, LBO = between groups of 8 along K.
// X is (Krows x NF) "row-major by 4-feature groups": addr(k, f) = (k/8)*LBO + (f/4)*SBO + (k%8)*16 + (f%4)*4.
// The kernel computes D[m][n] = sum_k A[k][m] * B[k][n] for M in {64, 128}, N = 64, K = 128 rows and
// dumps ALL 128 lanes x N columns so that the host can find where (m, n) lands for M = 64.
__device__ __forceinline__ uint32_t make_idesc_mn(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__global__ void __launch_bounds__(128) probe_mn(const float* __restrict__ A, const float* __restrict__ B,
                                                float* __restrict__ out, int M, int N, int KR, int swap) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* as = reinterpret_cast<float*>(smem);
  float* bs = as + KR * M;
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sboA = 128u, lboA = (uint32_t)(M / 4) * 128u;
  const uint32_t sboB = 128u, lboB = (uint32_t)(N / 4) * 128u;
  for (int i = tid; i < KR * M; i += 128) {
    const int k = i / M, f = i % M;
    as[((k / 8) * lboA + (f / 4) * sboA + (k % 8) * 16 + (f % 4) * 4) / 4] = A[i];
  }
  for (int i = tid; i < KR * N; i += 128) {
    const int k = i / N, f = i % N;
    bs[((k / 8) * lboB + (f / 4) * sboB + (k % 8) * 16 + (f % 4) * 4) / 4] = B[i];
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tbase_s)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tbase_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  {   // clear the accumulator columns so that untouched lanes read back as a sentinel
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = -777.f;
    for (int c = 0; c < N; c += 8) tc_st8(tbase + lane_base + c, v);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_mn(M, N, 1, 1);
    for (int kg = 0; kg < KR / 8; ++kg) {
      const uint64_t ad = swap ? make_bdesc(smem_u32(as) + kg * lboA, sboA, lboA)
                               : make_bdesc(smem_u32(as) + kg * lboA, lboA, sboA);
      const uint64_t bd = swap ? make_bdesc(smem_u32(bs) + kg * lboB, sboB, lboB)
                               : make_bdesc(smem_u32(bs) + kg * lboB, lboB, sboB);
      tc_mma_tf32_ss(tbase, ad, bd, idesc, kg > 0);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 8) {
    float r[8];
    tc_ld8(tbase + lane_base + c, r);
    for (int i = 0; i < 8; ++i) out[tid * N + c + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256)
                 : "memory");
}

static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  const int Ns[] = {64, 96, 160};
  const int Ks[] = {64, 56, 16};
  float *dA, *dB, *dO, *dR;
  long long* dC;
  CK(cudaMalloc(&dA, 128 * 64 * 4));
  CK(cudaMalloc(&dB, 256 * 64 * 4));
  CK(cudaMalloc(&dO, 128 * 256 * 4));
  CK(cudaMalloc(&dR, 128 * 32 * 4));
  CK(cudaMalloc(&dC, 8));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int mode = 0; mode < 3; ++mode)
    for (int ni = 0; ni < 3; ++ni)
      for (int ki = 0; ki < 3; ++ki)
        for (int swap = 0; swap < 1; ++swap) {
          const int N = Ns[ni], K = Ks[ki];
          std::vector<float> A(128 * K), W(N * K), Bc(N * K), O(128 * N), R(128 * 32);
          srand(1234 + N + K);
          for (auto& a : A)
            a = mode == 0 ? (float)((rand() % 17) - 8) / 8.f : (float)rand() / RAND_MAX * 2.f - 1.f;
          for (auto& w : W)
            w = mode == 0 ? (float)((rand() % 17) - 8) / 4.f : (float)rand() / RAND_MAX * 2.f - 1.f;
          // canonical no-swizzle K-major: [K/4][N][4]
          for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) Bc[(k / 4) * N * 4 + n * 4 + (k % 4)] = W[n * K + k];
          CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
          CK(cudaMemcpy(dB, Bc.data(), Bc.size() * 4, cudaMemcpyHostToDevice));
          CK(cudaMemset(dO, 0, 128 * 256 * 4));
          const int reps = mode == 2 ? 101 : 21;
          probe<<<1, 128, 2 * N * K * 4>>>(dA, dB, dO, dR, dC, N, K, mode, swap, reps);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
          CK(cudaMemcpy(R.data(), dR, R.size() * 4, cudaMemcpyDeviceToHost));
          long long cyc;
          CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
          double rterr = 0, err = 0, err1 = 0, ref_max = 0;
          for (int t = 0; t < 128; ++t)
            for (int c = 0; c < 32; ++c) rterr = fmax(rterr, fabs(R[t * 32 + c] - (t * 100 + c)));
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
              double s = 0, s1 = 0;
              for (int k = 0; k < K; ++k) {
                s += (double)A[m * K + k] * W[n * K + k];
                s1 += (double)tf32_trunc(A[m * K + k]) * tf32_trunc(W[n * K + k]);
              }
              err = fmax(err, fabs(O[m * N + n] - s));
              err1 = fmax(err1, fabs(O[m * N + n] - s1));
              ref_max = fmax(ref_max, fabs(s));
            }
          printf("mode=%d N=%3d K=%2d swap=%d  roundtrip_err=%g  max|D-fp64|=%.3e (vs 1xTF32 ref %.3e) "
                 "max|ref|=%.2f  cycles/layer=%lld\n",
                 mode, N, K, swap, rterr, err, err1, ref_max, cyc);
        }
  // ---- MN-major operands / M = 64 accumulator placement ----
  CK(cudaFuncSetAttribute(probe_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int M : {128, 64})
    for (int swap = 0; swap < 2; ++swap) {
      const int N = 64, KR = 128;
      std::vector<float> A(KR * M), B(KR * N), O(128 * N);
      srand(99);
      for (auto& a : A) a = (float)((rand() % 9) - 4) / 4.f;
      for (auto& b : B) b = (float)((rand() % 9) - 4) / 2.f;
      float *dA2, *dB2;
      CK(cudaMalloc(&dA2, A.size() * 4));
      CK(cudaMalloc(&dB2, B.size() * 4));
      CK(cudaMemcpy(dA2, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB2, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
      probe_mn<<<1, 128, (KR * M + KR * N) * 4>>>(dA2, dB2, dO, M, N, KR, swap);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mn M=%d swap=%d: CUDA error %s\n", M, swap, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
      // reference
      std::vector<double> R(M * N);
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double sacc = 0;
          for (int k = 0; k < KR; ++k) sacc += (double)A[k * M + m] * B[k * N + n];
          R[m * N + n] = sacc;
        }
      // hypothesis 1: row m in lane m
      double e1 = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) e1 = fmax(e1, fabs(O[m * N + n] - R[m * N + n]));
      // hypothesis 2 (M = 64): row m in lane (m % 16) + 32 * (m / 16)
      double e2 = 0;
      for (int m = 0; m < M && M == 64; ++m)
        for (int n = 0; n < N; ++n)
          e2 = fmax(e2, fabs(O[((m % 16) + 32 * (m / 16)) * N + n] - R[m * N + n]));
      int touched = 0;
      for (int l = 0; l < 128; ++l) touched += (O[l * N] != -777.f);
      printf("mn-major M=%3d swap=%d: lanes written %d; err(lane=m) %.3e; err(lane=m%%16+32*(m/16)) %.3e\n", M, swap,
             touched, e1, e2);
      if (M == 64) {
        printf("   lanes holding data:");
        for (int l = 0; l < 128; ++l) if (O[l * N] != -777.f) printf(" %d", l);
        printf("\n");
      }
      cudaFree(dA2); cudaFree(dB2);
    }
  CK(cudaFuncSetAttribute(cadence, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int var : {0, 4, 5})
    for (int N : {64, 128, 256}) {
      for (int nm : {48, 480}) {
        cadence<<<1, 128, (256 * 64 + 128 * 64) * 4>>>(dC, N, var, nm);
        CK(cudaDeviceSynchronize());
        long long cyc;
        CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
        printf("cadence var=%d N=%3d nmma=%3d: %lld cycles  (%.1f / MMA)\n", var, N, nm, cyc, (double)cyc / nm);
      }
    }
  return 0;
}
