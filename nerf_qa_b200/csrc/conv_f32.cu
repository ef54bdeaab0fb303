// CUDA-core convolutions: conv1_1 fused with the input normalisation (K = 27 is too thin for the
// tensor pipe and the layer is HBM-write bound), and the strict-fp32 3x3 convolution used by
// NQA_PREC_FP32 (parity path: fp32 operands, fp32 FMA accumulation like the reference's CPU conv).
//
// Reference semantics: (x-mean)/std then torchvision vgg16.features convs, 3x3 / stride 1 / zero pad 1 /
// bias / ReLU (nerf_qa/DISTS_pytorch/DISTS_pt.py:36-49, 91-103).
#include "common.cuh"
#include "kernels.h"

namespace nqa {

// ================================================================================================
// conv1_1: NCHW fp32 frames -> NHWC [n,h,w,64].  One thread per output pixel, 64 accumulators in
// registers, weights [27][64] broadcast from shared memory as float4.
// ================================================================================================
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(128)
conv1_1_kernel(const float* __restrict__ x, const float* __restrict__ y, int nx, int h, int w,
               const float* __restrict__ wpack, const float* __restrict__ bias, T* __restrict__ out,
               T* __restrict__ out_lo) {
  __shared__ __align__(16) float ws[27 * 64];
  __shared__ __align__(16) float bs[64];
  for (int i = threadIdx.x; i < 27 * 64; i += 128) ws[i] = wpack[i];
  if (threadIdx.x < 64) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();

  const int img = blockIdx.z, oy = blockIdx.y, ox = blockIdx.x * 128 + threadIdx.x;
  if (ox >= w) return;
  const float* src = img < nx ? x + (size_t)img * 3 * h * w : y + (size_t)(img - nx) * 3 * h * w;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};

  float in[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy - 1 + r;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = ox - 1 + s;
        float v = 0.f;  // zero padding applies to the NORMALISED tensor
        if (iy >= 0 && iy < h && ix >= 0 && ix < w)
          v = (__ldg(src + ((size_t)ci * h + iy) * w + ix) - mean[ci]) / stdv[ci];
        in[(ci * 3 + r) * 3 + s] = v;
      }
    }
  }

  float acc[64];
#pragma unroll
  for (int co = 0; co < 64; ++co) acc[co] = bs[co];
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const float v = in[k];
#pragma unroll
    for (int co = 0; co < 64; co += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(&ws[k * 64 + co]);
      acc[co + 0] = fmaf(v, wv.x, acc[co + 0]);
      acc[co + 1] = fmaf(v, wv.y, acc[co + 1]);
      acc[co + 2] = fmaf(v, wv.z, acc[co + 2]);
      acc[co + 3] = fmaf(v, wv.w, acc[co + 3]);
    }
  }
  constexpr int V = Vec<T>::N;
  const size_t o = (((size_t)img * h + oy) * w + ox) * 64;
#pragma unroll
  for (int co = 0; co < 64; co += V) {
    float v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = fmaxf(acc[co + i], 0.f);
    Vec<T>::store(out + o + co, v);
    if (SPLIT) {
      float l[V];
#pragma unroll
      for (int i = 0; i < V; ++i) l[i] = v[i] - to_float(from_float<T>(v[i]));
      Vec<T>::store(out_lo + o + co, l);
    }
  }
}

template <typename T>
int launch_conv1_1(const float* x, const float* y, int nx, int n, int h, int w, const float* wpack, const float* bias,
                   T* out, T* out_lo, cudaStream_t st) {
  dim3 grid(ceil_div(w, 128), h, n);
  if (out_lo)
    conv1_1_kernel<T, true><<<grid, 128, 0, st>>>(x, y, nx, h, w, wpack, bias, out, out_lo);
  else
    conv1_1_kernel<T, false><<<grid, 128, 0, st>>>(x, y, nx, h, w, wpack, bias, out, out_lo);
  return 1;
}
template int launch_conv1_1<float>(const float*, const float*, int, int, int, int, const float*, const float*, float*,
                                   float*, cudaStream_t);
template int launch_conv1_1<bf16>(const float*, const float*, int, int, int, int, const float*, const float*, bf16*,
                                  bf16*, cudaStream_t);

// ================================================================================================
// Strict fp32 3x3 convolution, NHWC.  Block = 8x16 output pixels x 64 output channels, 256 threads,
// each thread 8 pixels (one row segment) x 4 channels; input halo tile and a [9][16][64] weight slab
// staged in shared memory per 16-input-channel step.
// ================================================================================================
namespace f32conv {
constexpr int TH = 8, TW = 16, CK = 16, BN = 64;
constexpr int HALO = (TH + 2) * (TW + 2);  // 180 pixels
constexpr int IN_LD = CK + 1;
}  // namespace f32conv

__global__ void __launch_bounds__(256)
conv3x3_f32_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ wpack,
                   const float* __restrict__ bias, int h, int w, int cin, int cout, int tiles_w, int epi,
                   const float* __restrict__ mask) {
  using namespace f32conv;
  __shared__ float in_s[HALO * IN_LD];
  __shared__ __align__(16) float w_s[9 * CK * BN];

  const int img = blockIdx.z, n0 = blockIdx.y * BN;
  const int th = blockIdx.x / tiles_w, tw = blockIdx.x % tiles_w;
  const int h0 = th * TH, w0 = tw * TW;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row = ty >> 1, colbase = (ty & 1) * 8;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float* img_in = in + (size_t)img * h * w * cin;
  for (int c0 = 0; c0 < cin; c0 += CK) {
    __syncthreads();
    // halo tile: HALO pixels x CK channels, as float4s
    for (int i = threadIdx.x; i < HALO * (CK / 4); i += 256) {
      const int pix = i / (CK / 4), q = i % (CK / 4);
      const int iy = h0 - 1 + pix / (TW + 2), ix = w0 - 1 + pix % (TW + 2);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < h && ix >= 0 && ix < w)
        v = *reinterpret_cast<const float4*>(img_in + ((size_t)iy * w + ix) * cin + c0 + q * 4);
      float* d = &in_s[pix * IN_LD + q * 4];
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    // weight slab [9][CK][BN] from [9][cin][cout]
    for (int i = threadIdx.x; i < 9 * CK * (BN / 4); i += 256) {
      const int q = i % (BN / 4), kk = (i / (BN / 4)) % CK, tap = i / (BN / 4 * CK);
      const float4 v = *reinterpret_cast<const float4*>(wpack + ((size_t)tap * cin + c0 + kk) * cout + n0 + q * 4);
      *reinterpret_cast<float4*>(&w_s[(tap * CK + kk) * BN + q * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float* arow = &in_s[((row + r) * (TW + 2) + colbase) * IN_LD];
#pragma unroll 4
      for (int kk = 0; kk < CK; ++kk) {
        float a[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) a[i] = arow[i * IN_LD + kk];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const float4 bv = *reinterpret_cast<const float4*>(&w_s[((r * 3 + s) * CK + kk) * BN + tx * 4]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaf(a[i + s], bv.x, acc[i][0]);
            acc[i][1] = fmaf(a[i + s], bv.y, acc[i][1]);
            acc[i][2] = fmaf(a[i + s], bv.z, acc[i][2]);
            acc[i][3] = fmaf(a[i + s], bv.w, acc[i][3]);
          }
        }
      }
    }
  }
  const float4 bb = epi == kEpiBiasRelu ? *reinterpret_cast<const float4*>(bias + n0 + tx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int oy = h0 + row;
  if (oy < h) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ox = w0 + colbase + i;
      if (ox < w) {
        const size_t o = (((size_t)img * h + oy) * w + ox) * cout + n0 + tx * 4;
        float4 v;
        if (epi == kEpiBiasRelu) {
          v.x = fmaxf(acc[i][0] + bb.x, 0.f);
          v.y = fmaxf(acc[i][1] + bb.y, 0.f);
          v.z = fmaxf(acc[i][2] + bb.z, 0.f);
          v.w = fmaxf(acc[i][3] + bb.w, 0.f);
        } else {
          v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          if (epi == kEpiMask) {          // ReLU of the layer below
            const float4 a = *reinterpret_cast<const float4*>(mask + o);
            if (!(a.x > 0.f)) v.x = 0.f;
            if (!(a.y > 0.f)) v.y = 0.f;
            if (!(a.z > 0.f)) v.z = 0.f;
            if (!(a.w > 0.f)) v.w = 0.f;
          } else if (epi == kEpiDivide) { // gradient w.r.t. a pooled tensor, handed on as grad / pooled
            const float4 a = *reinterpret_cast<const float4*>(mask + o);
            v.x /= a.x; v.y /= a.y; v.z /= a.z; v.w /= a.w;
          }
        }
        *reinterpret_cast<float4*>(out + o) = v;
      }
    }
  }
}

int launch_conv3x3_f32(const float* in, float* out, const float* wpack, const float* bias, int n, int h, int w, int cin,
                       int cout, cudaStream_t st, int epi, const float* mask) {
  using namespace f32conv;
  const int tiles_w = ceil_div(w, TW), tiles_h = ceil_div(h, TH);
  dim3 grid(tiles_w * tiles_h, cout / BN, n);
  conv3x3_f32_kernel<<<grid, 256, 0, st>>>(in, out, wpack, bias, h, w, cin, cout, tiles_w, epi, mask);
  return 1;
}

}  // namespace nqa
