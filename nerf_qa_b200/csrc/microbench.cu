// Measured denominators for the CUDA-core and warp-level-MMA rooflines (bench.py): MEASURED_PEAKS.json has the HBM
// copy and the cuBLAS bf16 (tcgen05) figures only, and the A-DISTS window kernels are bound by neither.
//   fp32_fma  : packed f32x2 FMAs (FFMA2, what the window kernel issues), 8 independent chains per thread;
//   mma_sync  : warp-level mma.sync.m16n8k16 bf16 -> fp32 (HMMA), 8 independent accumulators per warp -- the
//               instruction the banded-Toeplitz window kernel (adists_tc.cu) is built on.
// Both run `iters` dependent rounds on every SM and are timed with CUDA events by the caller-facing entry point.
#include "common.cuh"
#include "kernels.h"

namespace nqa {
namespace {

__global__ void __launch_bounds__(256) fma2_kernel(int iters, float seed, float* __restrict__ out) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(seed + i, seed - i);
  const float2 a = make_float2(1.0000001f, 0.9999999f), b = make_float2(seed * 1e-9f, -seed * 1e-9f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(a, acc[i], b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  if (s == 123.456f) out[0] = s;   // never true: keeps the chains alive
}

__device__ inline void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256) mma_sync_kernel(int iters, uint32_t seed, float* __restrict__ out) {
  float d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a[4], b[2];
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = 0x3c003c00u + seed + threadIdx.x + j;   // small finite bf16 pairs
  b[0] = 0x3a803a80u + seed;
  b[1] = 0x3a003a00u + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) mma_bf16_16816(d[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456f) out[0] = s;
}

}  // namespace

// Returns the achieved rate in FMA/s (fp32_fma: scalar FMAs, two per FFMA2 lane; mma_sync: 16*8*16 per instruction).
int microbench_rate(int which, int sm_count, float* scratch, cudaStream_t st, double* fma_per_s) {
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return 1;
  const int blocks = sm_count * 8;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    const int iters = 4096;
    cudaEventRecord(e0, st);
    if (which == 0)
      fma2_kernel<<<blocks, 256, 0, st>>>(iters, 1.5f, scratch);
    else
      mma_sync_kernel<<<blocks, 256, 0, st>>>(iters, 0u, scratch);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return 1; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double per_thread = which == 0 ? (double)iters * 4 * 8 * 2 : 0.0;
    const double per_warp = which == 0 ? 0.0 : (double)iters * 4 * 8 * (16.0 * 8 * 16);
    const double total = which == 0 ? per_thread * 256.0 * blocks : per_warp * 8.0 * blocks;
    const double rate = total / ((double)ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;   // first repetition is the warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *fma_per_s = best;
  return 0;
}

}  // namespace nqa
