// Accept / reject + order-preserving compaction of one batch of rejection-sampling proposals
// (/root/reference/sbi/samplers/rejection/rejection.py:170-200: `target_proposal_ratio = exp(potential -
// scaled proposal log-prob); keep = ratio > u; samples = candidates[keep]`): the accepted candidates are
// written, in proposal order, behind the ones already collected, together with their global proposal
// index; the running count stays on the device.  Three launches (flags + block counts, scan of the block
// counts, scatter) replace exp / compare / boolean-index (whose `nonzero` synchronises the host) -- the
// accepted set and its order are exactly those of the reference's boolean indexing.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/sbi_b200.h"
#include "device.cuh"

namespace sbi {
namespace compact {

constexpr int kThreads = 256;
constexpr int kPer = 4;                       // consecutive candidates per thread
constexpr int kTile = kThreads * kPer;        // candidates per block

__device__ __forceinline__ bool keep_flag(const float* lt, const float* ls, const float* u, int64_t i) {
  // same expression as the reference: exp(a - b) > u  (NaN compares false -> rejected)
  return expf(lt[i] - ls[i]) > u[i];
}

__global__ void count_kernel(const float* __restrict__ lt, const float* __restrict__ ls,
                             const float* __restrict__ u, int64_t n, int32_t* __restrict__ block_count) {
  __shared__ int sh[kThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kPer;
  int c = 0;
#pragma unroll
  for (int j = 0; j < kPer; ++j)
    if (base + j < n && keep_flag(lt, ls, u, base + j)) ++c;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
    block_count[blockIdx.x] = t;
  }
}

// exclusive scan of the block counts (in place), shifted by the number already collected; *count += total
__global__ void scan_kernel(int32_t* __restrict__ block_count, int nb, int32_t* __restrict__ count) {
  __shared__ int sh[kThreads];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = *count;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += kThreads) {
    const int i = b0 + threadIdx.x;
    const int v = i < nb ? block_count[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kThreads; o <<= 1) {          // Hillis-Steele inclusive scan
      const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) block_count[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[kThreads - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

__global__ void scatter_kernel(const float* __restrict__ cand, int D, const float* __restrict__ lt,
                               const float* __restrict__ ls, const float* __restrict__ u, int64_t n,
                               int64_t index_base, const int32_t* __restrict__ block_off, float* __restrict__ out,
                               int64_t* __restrict__ out_idx, int64_t cap) {
  __shared__ int sh[kThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kPer;
  bool k[kPer];
  int c = 0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    k[j] = base + j < n && keep_flag(lt, ls, u, base + j);
    c += k[j] ? 1 : 0;
  }
  // exclusive prefix of c over the block (threads own consecutive candidates, so thread order = proposal order)
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += sh[w];
  int64_t pos = (int64_t)block_off[blockIdx.x] + woff + incl - c;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    if (!k[j]) continue;
    if (pos < cap) {
      for (int d = 0; d < D; ++d) out[pos * D + d] = cand[(base + j) * D + d];
      if (out_idx != nullptr) out_idx[pos] = index_base + base + j;
    }
    ++pos;
  }
}

}  // namespace compact
}  // namespace sbi

using namespace sbi;

extern "C" int64_t sbi_b200_reject_scratch_ints(int64_t n) {
  return n < 1 ? 1 : (n + compact::kTile - 1) / compact::kTile;
}

extern "C" int sbi_b200_reject_compact(const float* d_cand, int32_t D, const float* d_log_target,
                                       const float* d_log_scaled, const float* d_u, int64_t n, int64_t index_base,
                                       float* d_out, int64_t* d_out_idx, int64_t cap, int32_t* d_count,
                                       int32_t* d_scratch, void* stream) {
  sbi::DeviceGuard dev_guard_(d_cand);
  if (!d_cand || !d_log_target || !d_log_scaled || !d_u || !d_out || !d_count || !d_scratch || D < 1 || n < 0 ||
      cap < 0)
    return SBI_EINVAL;
  if (n == 0) return 0;
  const int64_t nb64 = (n + compact::kTile - 1) / compact::kTile;
  if (nb64 > (1 << 30)) return SBI_EINVAL;
  const int nb = (int)nb64;
  cudaStream_t s = (cudaStream_t)stream;
  compact::count_kernel<<<nb, compact::kThreads, 0, s>>>(d_log_target, d_log_scaled, d_u, n, d_scratch);
  compact::scan_kernel<<<1, compact::kThreads, 0, s>>>(d_scratch, nb, d_count);
  compact::scatter_kernel<<<nb, compact::kThreads, 0, s>>>(d_cand, D, d_log_target, d_log_scaled, d_u, n, index_base,
                                                         d_scratch, d_out, d_out_idx, cap);
  return (int)cudaGetLastError();
}
