// Micro-benchmark: issue throughput of legacy mma.sync (TF32 m16n8k8, BF16 m16n8k16) vs FFMA on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tput mma_tput.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__global__ void k_tf32(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = b0 + 1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = b0 + 1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters) {
  float c[32];
  for (int i = 0; i < 32; ++i) c[i] = i;
  float a = threadIdx.x * 1e-3f, b = 1.0001f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = fmaf(c[i], b, a);
  }
  float s = 0; for (int i = 0; i < 32; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K> float run(K k, int warps, int iters, float* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<148, warps * 32>>>(d, 10);
  cudaEventRecord(e0); k<<<148, warps * 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 20000;
  for (int warps : {1, 2, 4, 8, 16}) {
    float t1 = run(k_tf32, warps, iters, d), t2 = run(k_bf16, warps, iters, d), t3 = run(k_ffma, warps, iters, d);
    double mma = 148.0 * warps * iters * 8;
    printf("warps/SM %2d: tf32 m16n8k8 %.1f TFLOP/s (%.2f mma/clk/SM @1.9GHz) | bf16 m16n8k16 %.1f TFLOP/s | ffma %.1f TFLOP/s\n",
           warps, mma * 2 * 16 * 8 * 8 / t1 / 1e9, mma / 148 / (t1 * 1e-3 * 1.9e9), mma * 2 * 16 * 8 * 16 / t2 / 1e9,
           148.0 * warps * 32 * iters * 32 * 2 / t3 / 1e9);
  }
  return 0;
}
