// Building blocks of the tensor-core (tcgen05) row-tile kernels: PTX wrappers, the 3xTF32 split,
// the weight-stage ring and its issuing logic.  Used by nsf_tc.cu and ratio_tc.cu; see the header of
// nsf_tc.cu for the design.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sbi_b200.h"
#include "common.cuh"

namespace sbi {
namespace tc {

// Phase timeline of CTA 0 (tuning builds only: -DSBI_TC_TIMELINE, profiles/tc_timeline.py): thread 0
// appends (id, clock64) pairs to a global buffer read back through sbi_b200_debug_timeline_{fwd,bwd}().
#ifdef SBI_TC_TIMELINE
static __device__ unsigned long long g_tl[4096];     // one copy per translation unit
static __device__ int g_tl_n;
#define SBI_TL(id)                                                                        \
  do {                                                                                    \
    if (blockIdx.x == 0 && threadIdx.x == 0 && g_tl_n < 2047) {                           \
      g_tl[2 * g_tl_n] = (unsigned long long)(id);                                        \
      g_tl[2 * g_tl_n + 1] = clock64();                                                   \
      ++g_tl_n;                                                                           \
    }                                                                                     \
  } while (0)
#else
#define SBI_TL(id) do { } while (0)
#endif

constexpr int kRows = 128;        // rows per tile
constexpr int kRowThreads = 256;  // two threads per row (column halves)
constexpr int kThreads = 256;     // 8 row warps, which also take turns issuing MMAs and TMA copies
constexpr int kLuMax = 16;        // LULinear runs on register-resident rows of <= 16 features
constexpr int kSlots = 3;         // weight ring: up to two stages in use + one prefetched
constexpr int kCols = 256;        // TMEM columns per CTA
constexpr int cAhi = 0, cAlo = 64, cD = 128, cG = 192;

// ---- tcgen05 wrappers -------------------------------------------------------------------------
__device__ __forceinline__ void fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : one K = 8 step of kind::tf32
__device__ __forceinline__ void mma_tf32(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): fp32 accumulate @4,
// A/B format tf32 @7/@10, both K-major, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// shared-memory operand descriptor, SWIZZLE_NONE, K-major: core matrix = 8 rows x 16 B contiguous;
// LBO = byte distance of K-adjacent core matrices, SBO = byte distance of 8-row groups
// (field positions: cute/arch/mma_sm100_desc.hpp SmemDescriptor; the assignment was pinned on
// hardware with profiles/micro/umma_probe.cu)
__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}
// 8 consecutive columns of the thread's lane -> v[0..8)   (no wait inside)
__device__ __forceinline__ void ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st4(uint32_t taddr, const float (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3]))
               : "memory");
}
__device__ __forceinline__ void ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
template <int NCHUNK>
__device__ __forceinline__ void ld_cols(uint32_t taddr, float* v) {
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) ld8(taddr + 8 * c, v + 8 * c);
}

// hi = x rounded to tf32 (10 explicit mantissa bits, round half away in the integer domain),
// lo = x - hi (exact in fp32).  cvt.rna.tf32.f32 computes the same hi but expands to a longer
// sequence on sm_100a.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}
// split 8 values and put them into A_hi / A_lo columns [col, col+8) of the thread's lane
__device__ __forceinline__ void store_a8(uint32_t tlane, int col, const float (&v)[8]) {
  float hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_tf32(v[i], hi[i], lo[i]);
  st8(tlane + cAhi + col, hi);
  st8(tlane + cAlo + col, lo);
}
__device__ __forceinline__ void store_a4(uint32_t tlane, int col, const float (&v)[4]) {
  float hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(v[i], hi[i], lo[i]);
  st4(tlane + cAhi + col, hi);
  st4(tlane + cAlo + col, lo);
}
__device__ __forceinline__ void group_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---- weight re-pack: flat fp32 parameters -> [hi | lo] UMMA operand blocks ---------------------
static __global__ void tc_pack_kernel(const float* __restrict__ params, const int32_t* __restrict__ src,
                                   float* __restrict__ tcw, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = __ldg(src + i);
  float v = 0.f;
  if (s >= 0) {
    float hi, lo;
    split_tf32(__ldg(params + s), hi, lo);
    v = hi;
  } else if (s <= -2) {
    float hi, lo;
    split_tf32(__ldg(params + (-2 - s)), hi, lo);
    v = lo;
  }
  tcw[i] = v;
}

// ---- the kernel ------------------------------------------------------------------------------------
// The warps of the CTA take turns driving the tensor core and the weight stream (stage k is
// issued by warp k % 8, so the issue work is spread evenly): the whole warp runs this code
// converged (so the descriptor arithmetic stays in the uniform datapath and the MMAs go out at
// the tensor pipe's own cadence; a single divergent thread issues 2x slower, see
// profiles/micro/umma_probe.cu), one elected lane executes the tcgen05 / TMA instructions.
// Stage k lives in ring slot
// k % kSlots.  A stage is fetched (TMA bulk copy, completion on full[slot]) as soon as the stage
// that used its slot kSlots stages earlier is known to be complete, which every warp learns each
// time it passes an accumulator barrier (a tcgen05.commit covers every MMA issued before it).
template <int NSLOT>
struct IssuerT {
  uint32_t tbase;       // TMEM base (lane 0, column 0)
  bool leader;          // the elected lane of this warp
  int warp;             // this warp; stage k is issued by warp k % 8, fetched by warp (k+4) % 8
  bool mine;            // this warp issues the current stage
  float* ring;
  uint64_t *full, *bars;
  const float* tcw;
  const int32_t* tab;   // stage table (all layers)
  int cap, T;
  uint32_t it;          // stages issued
  uint32_t done;        // stages known complete
  uint32_t fetched;     // stages fetched
  uint32_t cov0, cov1;  // stages covered by the last commit on each accumulator barrier
  uint32_t sbase, lo_off;   // current stage: shared address of the hi half, byte offset of lo half
  int64_t f_tile, ntiles, tile_step;   // next stage to fetch
  int f_l, f_s;
  bool reverse;         // layers are walked T-1 .. 0 (sampling direction)

  __device__ __forceinline__ void pump() {
    while (fetched < done + NSLOT && f_tile < ntiles) {
      const int32_t* t = tab + (reverse ? T - 1 - f_l : f_l) * SBI_NSF_TC_STRIDE;
      const int off = __ldg(t + 4 + 4 * f_s), nfl = __ldg(t + 5 + 4 * f_s);
      const uint32_t slot = fetched % NSLOT;
      if (leader && (int)((fetched + 4u) & 7u) == warp) {
        mbar_arrive_expect_tx(&full[slot], (uint32_t)nfl * 4u);
        bulk_g2s(ring + (size_t)slot * cap, tcw + off, (uint32_t)nfl * 4u, &full[slot]);
      }
      ++fetched;
      if (++f_s == __ldg(t)) {
        f_s = 0;
        if (++f_l == T) { f_l = 0; f_tile += tile_step; }
      }
    }
  }
  __device__ __forceinline__ void begin(int stage_floats) {
    mine = (int)(it & 7u) == warp;
    if (!mine) return;
    const uint32_t s = it % NSLOT;
    mbar_wait(&full[s], (it / NSLOT) & 1u);
    // (the shuffles only tell the compiler that these values are warp-uniform)
    sbase = __shfl_sync(0xffffffffu, smem_u32(ring + (size_t)s * cap), 0);
    lo_off = __shfl_sync(0xffffffffu, (uint32_t)stage_floats * 2u, 0);   // (floats / 2) * 4 bytes
    fence_after();
  }
  // one operand block of N rows starting `blk_floats` into the half: nk K-steps, A columns from a0
  __device__ __forceinline__ void block(int dcol, int a0, int nk, int blk_floats, int N, uint32_t& acc) {
    if (!mine) return;
    N = __shfl_sync(0xffffffffu, N, 0);
    nk = __shfl_sync(0xffffffffu, nk, 0);
    blk_floats = __shfl_sync(0xffffffffu, blk_floats, 0);
    const uint32_t idesc = make_idesc(N);
    const uint32_t slab = (uint32_t)N * 16u;
    const uint32_t bh = sbase + (uint32_t)blk_floats * 4u;
    uint64_t dh = make_bdesc(bh, slab, 128u);
    uint64_t dl = make_bdesc(bh + lo_off, slab, 128u);
    const uint64_t dstep = (uint64_t)((2u * slab) >> 4);    // start-address field advance per K-step
    uint32_t ah = tbase + cAhi + a0, al = tbase + cAlo + a0;
    const uint32_t d = tbase + dcol;
#pragma unroll 8
    for (int kk = 0; kk < nk; ++kk) {
      if (leader) {
        mma_tf32(d, ah, dh, idesc, acc);
        mma_tf32(d, al, dh, idesc, 1u);
        mma_tf32(d, ah, dl, idesc, 1u);
      }
      acc = 1u;
      dh += dstep; dl += dstep; ah += 8; al += 8;
    }
  }
  // close the stage: its accumulators are signalled on accumulator barrier `b`
  __device__ __forceinline__ void end(int b) {
    if (mine && leader) commit(&bars[b]);
    ++it;
    if (b == 0) cov0 = it; else cov1 = it;
  }
  // the warp has just passed accumulator barrier b
  __device__ __forceinline__ void passed(int b) {
    const uint32_t c = (b == 0) ? cov0 : cov1;
    if (c > done) done = c;
    pump();
  }
};
using Issuer = IssuerT<kSlots>;

// instruction descriptor with an explicit M (64 or 128)
__device__ __forceinline__ uint32_t make_idesc_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T : one K = 8 step of kind::tf32, both operands from shared memory
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// TMA bulk copy shared -> global (bulk async-group of the calling thread); bytes % 16 == 0, both
// addresses 16-byte aligned.  `add`: element-wise fp32 reduction into global instead of a plain store.
__device__ __forceinline__ void bulk_s2g(float* dst_gmem, const float* src_smem, uint32_t bytes, bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
  else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the committed groups have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// the committed groups are complete (their global writes are performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tc
}  // namespace sbi
