// Shared device/host helpers for the nerfqa_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <utility>

namespace nqa {

constexpr int kNumLevels = 6;
constexpr int kNumConvs = 13;
// 1 / std of the ImageNet normalisation (DISTS_pt.py:52, 92), correctly rounded: the bf16 trunks multiply by the
// reciprocal (the fused and the stand-alone conv1_1 kernels use the same constants and stay bit-identical); the fp32
// trunk divides like the reference.
constexpr float kInvStd0 = 1.f / 0.229f, kInvStd1 = 1.f / 0.224f, kInvStd2 = 1.f / 0.225f;
constexpr int kTotalChannels = 1475;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ constexpr int level_channels(int k) {
  return k == 0 ? 3 : k == 1 ? 64 : k == 2 ? 128 : k == 3 ? 256 : 512;
}
__host__ __device__ constexpr int level_offset(int k) {
  return k == 0 ? 0 : k == 1 ? 3 : k == 2 ? 67 : k == 3 ? 195 : k == 4 ? 451 : 963;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE and per KERNEL: remember the largest request per
// (kernel entry point, device) so that a process holding contexts on several GPUs configures each of them.
inline void ensure_dynamic_smem_impl(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> granted;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  int& g = granted[std::make_pair(kernel, dev)];
  if (bytes > g) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    g = bytes;
  }
}
template <typename K>
inline void ensure_dynamic_smem(K kernel, int bytes) {
  ensure_dynamic_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}

__host__ __device__ inline int pooled(int n) { return (n - 1) / 2 + 1; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- activation element types: fp32 (strict) or bf16 (tensor path), 16-byte vectors -------------
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static inline void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static inline void load_stream(const float* p, float (&v)[4]) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static inline void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static inline void unpack(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static inline void load(const __nv_bfloat16* p, float (&v)[8]) {
    unpack(*reinterpret_cast<const uint4*>(p), v);
  }
  __device__ static inline void load_stream(const __nv_bfloat16* p, float (&v)[8]) {
    unpack(__ldcs(reinterpret_cast<const uint4*>(p)), v);
  }
  __device__ static inline uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  __device__ static inline void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    t.x = pack2(v[0], v[1]); t.y = pack2(v[2], v[3]); t.z = pack2(v[4], v[5]); t.w = pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

__device__ inline float to_float(float v) { return v; }
__device__ inline float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ inline T from_float(float v);
template <> __device__ inline float from_float<float>(float v) { return v; }
template <> __device__ inline __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ inline float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in every thread.
template <typename T> __device__ inline T block_sum(T v, T* smem /* >= 33 elements */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  if (wid == 0) {
    T t = lane < nw ? smem[lane] : T(0);
    t = warp_sum(t);
    if (lane == 0) smem[32] = t;
  }
  __syncthreads();
  return smem[32];
}

}  // namespace nqa
