// Data-parallel gradient exchange over NVLink peer memory, fused with the reduction that the
// optimizer needs: replaces [NCCL all-reduce of the flat gradient -> norm pass] of the multi-GPU
// training step (reference semantics: every rank applies clip_grad_norm_ + Adam to the
// SUM of the per-rank gradients, /root/reference/sbi/inference/trainers/base.py:1181-1187).
//
// Every rank owns a "symmetric" buffer (plain cudaMalloc, exported with cudaIpcGetMemHandle and
// mapped by the other ranks of the node):
//     [ grad slot 0 | grad slot 1 | flags slot 0 | flags slot 1 | err ]
// One kernel per step and rank, block b owning 256 consecutive gradient entries:
//   1. publish: copy its slice of the local gradient into the rank's own slot (step parity picks
//      the slot), __threadfence_system(), flag[slot][b] = step + 1;
//   2. wait until flag[slot][b] of every peer equals step + 1 (P2P loads over NVLink; bounded spin;
//      equality, so that flags left by a rewound step counter can never satisfy it early);
//   3. sum the slice over the ranks IN RANK ORDER from the symmetric buffers (own rank included, so
//      every rank adds the same numbers in the same order and the replicas stay bit-identical),
//      write the reduced gradient and one partial of sum(g^2) per block for the clip norm.
// Block b only ever waits for block b of the other ranks, so no co-residency is required.  The step
// number is a device counter -- the exchange's OWN, kept in the symmetric buffer and advanced by the
// last block to finish (d_step == nullptr; it only ever grows, so flags left by a warm-up pass whose
// optimizer state was rewound can never match a later step), or a caller-provided one (the
// optimizer's) -- so the launch is CUDA-graph capturable and carries no host-side state.  Slot reuse is safe: a rank rewrites slot s two steps later, after it has seen
// every peer's flag of the step in between, which that peer set after finishing this step's reads.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/sbi_b200.h"
#include "common.cuh"
#include "device.cuh"

namespace sbi {

constexpr int kPeerBlock = 256;          // threads = gradient entries per block (64 float4 columns x 4)
constexpr int kMaxPeers = 16;

struct PeerPtrs {
  float* p[kMaxPeers];
};

__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kPeerBlock)
peer_sum_kernel(const float* __restrict__ grad_local, PeerPtrs peers, int world, int rank, int64_t n,
                int64_t n_pad, int nblk, float* __restrict__ grad_out, const uint8_t* __restrict__ mask,
                float* __restrict__ sumsq_part, const int32_t* __restrict__ d_step) {
  __shared__ float red[32];
  __shared__ int s_bad;
  const int64_t i = (int64_t)blockIdx.x * kPeerBlock + threadIdx.x;
  float* own = peers.p[rank];
  int* ctl = reinterpret_cast<int*>(own + 2 * n_pad) + 2 * nblk;     // [err | done | counter | -]
  // this exchange, 1-based (every block reads the counter before the last block can advance it)
  const int step = (d_step ? __ldg(d_step) : *reinterpret_cast<volatile int*>(ctl + 2)) + 1;
  const int slot = step & 1;
  // 1. publish
  own[(int64_t)slot * n_pad + i] = (i < n) ? grad_local[i] : 0.f;
  if (threadIdx.x == 0) s_bad = 0;
  __threadfence_system();
  __syncthreads();
  int* own_flags = reinterpret_cast<int*>(own + 2 * n_pad);
  if (threadIdx.x == 0) st_flag(own_flags + slot * nblk + blockIdx.x, step);
  // 2. wait for block b of every peer (bounded: ~4 s, then flag the error and go on)
  if (threadIdx.x < world && threadIdx.x != rank) {
    const int* f = reinterpret_cast<const int*>(peers.p[threadIdx.x] + 2 * n_pad) + slot * nblk + blockIdx.x;
    const long long t0 = clock64();
    while (ld_flag(f) != step) {
      if (clock64() - t0 > 8000000000LL) { s_bad = 1; break; }
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (s_bad && threadIdx.x == 0) reinterpret_cast<int*>(own + 2 * n_pad)[2 * nblk] = 1;
  // 3. sum over ranks in rank order
  float a = 0.f;
  for (int r = 0; r < world; ++r) a += __ldcv(peers.p[r] + (int64_t)slot * n_pad + i);
  float ss = 0.f;
  if (i < n) {
    grad_out[i] = a;
    if (mask == nullptr || mask[i]) ss = a * a;
  }
  if (sumsq_part != nullptr) {
    ss = warp_sum(ss);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = ss;
    __syncthreads();
    if (w == 0) {
      float t = (l < kPeerBlock / 32) ? red[l] : 0.f;
      t = warp_sum(t);
      if (l == 0) sumsq_part[blockIdx.x] = t;
    }
  }
  // own counter: the block that finishes last advances it (all blocks have read it by then)
  if (d_step == nullptr && threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ctl + 1, 1) == nblk - 1) {
      ctl[1] = 0;
      ctl[2] = step;
      __threadfence();
    }
  }
}

}  // namespace sbi

using namespace sbi;

extern "C" int64_t sbi_b200_peer_bytes(int64_t n_params) {
  const int64_t nblk = (n_params + kPeerBlock - 1) / kPeerBlock;
  return (2 * nblk * kPeerBlock) * 4 + (2 * nblk + 4) * 4;
}
extern "C" int sbi_b200_peer_blocks(int64_t n_params) { return (int)((n_params + kPeerBlock - 1) / kPeerBlock); }

extern "C" void* sbi_b200_peer_alloc(int64_t n_params) {
  void* p = nullptr;
  const int64_t bytes = sbi_b200_peer_bytes(n_params);
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  if (cudaMemset(p, 0, bytes) != cudaSuccess) return nullptr;
  cudaDeviceSynchronize();
  return p;
}
extern "C" int sbi_b200_peer_free(void* p) {
  sbi::DeviceGuard dev_guard_(p); return (int)cudaFree(p); }
extern "C" int sbi_b200_peer_export(void* p, void* handle64) {
  sbi::DeviceGuard dev_guard_(p);
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
extern "C" void* sbi_b200_peer_import(const void* handle64) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
extern "C" int sbi_b200_peer_close(void* p) {
  sbi::DeviceGuard dev_guard_(p); return (int)cudaIpcCloseMemHandle(p); }

extern "C" int sbi_b200_peer_sum(const float* d_grad_local, void* const* h_peer_ptrs, int world, int rank,
                                 int64_t n_params, float* d_grad_out, const uint8_t* d_mask,
                                 float* d_sumsq_part, const int32_t* d_step, void* stream) {
  sbi::DeviceGuard dev_guard_(d_grad_local);
  if (!d_grad_local || !h_peer_ptrs || world < 1 || world > kMaxPeers || rank < 0 || rank >= world ||
      n_params < 1 || !d_grad_out)
    return SBI_EINVAL;
  PeerPtrs pp;
  for (int r = 0; r < kMaxPeers; ++r) pp.p[r] = r < world ? static_cast<float*>(h_peer_ptrs[r]) : nullptr;
  for (int r = 0; r < world; ++r)
    if (!pp.p[r]) return SBI_EINVAL;
  const int nblk = sbi_b200_peer_blocks(n_params);
  peer_sum_kernel<<<nblk, kPeerBlock, 0, (cudaStream_t)stream>>>(d_grad_local, pp, world, rank, n_params,
                                                               (int64_t)nblk * kPeerBlock, nblk, d_grad_out,
                                                               d_mask, d_sumsq_part, d_step);
  return (int)cudaGetLastError();
}

/* 1 if a bounded wait of the peer kernel expired since the buffer was allocated */
extern "C" int sbi_b200_peer_error(const void* p, int64_t n_params) {
  sbi::DeviceGuard dev_guard_(p);
  const int nblk = sbi_b200_peer_blocks(n_params);
  int v = 0;
  const int* f = reinterpret_cast<const int*>(static_cast<const float*>(p) + 2 * (int64_t)nblk * kPeerBlock) + 2 * nblk;
  if (cudaMemcpy(&v, f, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}
