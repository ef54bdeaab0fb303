// Backward pass of the DISTS score w.r.t. the frames / features (row a9 of SURVEY.md §8: DISTS.forward with
// require_grad=True, DISTS_pt.py:105-148 under autograd; forward_from_feats, DISTS_pt.py:181-208).
//
//   score_b = 1 - sum_k sum_c (alpha_c S1 + beta_c S2) / w,   S1 = (2 mx my + c1)/(mx^2 + my^2 + c1),
//   S2 = (2 cov + c2)/(vx + vy + c2),  mx = E[x], vx = E[x^2] - mx^2, cov = E[xy] - mx my  (per level, image, channel)
//
// so the gradient w.r.t. one feature element is AFFINE in the element pair:
//   d score / d x_i = Px + Qx x_i + R y_i,      d score / d y_i = Py + Qy y_i + R x_i
// with five coefficients per (image pair, channel) that depend only on the saved moments (stat_coef_kernel).
// The trunk is differentiated layer by layer: the data gradient of a 3x3 convolution is the same convolution with
// rotated / transposed weights (the tcgen05 kernels with a masking epilogue, kernels.h ConvEpilogue), L2pooling
// out = sqrt(sum g x^2 + eps) has d out / d x_j = g x_j / out, and conv1_1 (3 input channels) is differentiated on
// the CUDA cores together with the input normalisation.
#include "common.cuh"
#include "kernels.h"

namespace nqa {

// ================================================================================================
// coefficients: grid = (ceil(1475/256), pairs)
// ================================================================================================
__global__ void __launch_bounds__(256)
stat_coef_kernel(const float* __restrict__ moments, const float* __restrict__ alpha, const float* __restrict__ beta,
                 int normalize, const float* __restrict__ grad_score, GradLevels lv, int levels_valid,
                 float* __restrict__ coef) {
  __shared__ double smd[33];
  double wpart = 0.0;
  for (int c = threadIdx.x; c < kTotalChannels; c += 256) wpart += (double)alpha[c] + (double)beta[c];
  double wsum = block_sum<double>(wpart, smd);
  if (!normalize) wsum = 1.0;          // channel weights are used as given (nqa_set_channel_weights)
  const int b = blockIdx.y;
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= kTotalChannels) return;
  const int k = c < 3 ? 0 : c < 67 ? 1 : c < 195 ? 2 : c < 451 ? 3 : c < 963 ? 4 : 5;
  float* out = coef + (size_t)b * 5 * kCoefStride + coef_offset(k) + (c - level_offset(k));   // out[q * kCoefStride]
  if (k >= levels_valid) {
#pragma unroll
    for (int q = 0; q < 5; ++q) out[q * kCoefStride] = 0.f;
    return;
  }
  const float* m = moments + ((size_t)b * kTotalChannels + c) * 5;
  const double mx = m[0], my = m[1], vx = m[2], vy = m[3], cov = m[4];
  const double g = grad_score[b];
  const double a = -g * (double)alpha[c] / wsum, bb = -g * (double)beta[c] / wsum;
  const double c1 = 1e-6, c2 = 1e-6, inv_n = lv.inv_n[k];
  const double n1 = 2.0 * mx * my + c1, d1 = mx * mx + my * my + c1;
  const double ds1_dmx = (2.0 * my * d1 - n1 * 2.0 * mx) / (d1 * d1);
  const double ds1_dmy = (2.0 * mx * d1 - n1 * 2.0 * my) / (d1 * d1);
  const double n2 = 2.0 * cov + c2, d2 = vx + vy + c2;
  const double ds2_dcov = 2.0 / d2, ds2_dv = -n2 / (d2 * d2);
  out[0 * kCoefStride] = (float)((a * ds1_dmx - 2.0 * bb * ds2_dv * mx - bb * ds2_dcov * my) * inv_n);   // Px
  out[1 * kCoefStride] = (float)(2.0 * bb * ds2_dv * inv_n);                                             // Qx (= Qy)
  out[2 * kCoefStride] = (float)(bb * ds2_dcov * inv_n);                                                 // R
  out[3 * kCoefStride] = (float)((a * ds1_dmy - 2.0 * bb * ds2_dv * my - bb * ds2_dcov * mx) * inv_n);   // Py
  out[4 * kCoefStride] = out[1 * kCoefStride];                                                           // Qy
}

int launch_stat_coef(const float* moments, const float* alpha, const float* beta, bool normalize, const float* grad_score,
                     const GradLevels& lv, int levels_valid, int pairs, float* coef, cudaStream_t st) {
  dim3 grid(ceil_div(kTotalChannels, 256), pairs);
  stat_coef_kernel<<<grid, 256, 0, st>>>(moments, alpha, beta, normalize ? 1 : 0, grad_score, lv, levels_valid, coef);
  return 1;
}

// ================================================================================================
// planar (NCHW fp32) statistics gradient: level 0 and every level of forward_from_feats.
// grid = (ceil(hw/1024), C, pairs)
// ================================================================================================
__global__ void __launch_bounds__(256)
stat_grad_planar_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ coef,
                        int level, int c, int hw, float* __restrict__ gx, float* __restrict__ gy, int accumulate) {
  const int ch = blockIdx.y, b = blockIdx.z;
  const float* k = coef + (size_t)b * 5 * kCoefStride + coef_offset(level) + ch;
  const float px = k[0], qx = k[kCoefStride], r = k[2 * kCoefStride], py = k[3 * kCoefStride], qy = k[4 * kCoefStride];
  const size_t base = ((size_t)b * c + ch) * hw;
  for (int i = blockIdx.x * 1024 + threadIdx.x; i < min(hw, (int)(blockIdx.x + 1) * 1024); i += 256) {
    const float a = x[base + i], d = y[base + i];
    const float vx = fmaf(qx, a, fmaf(r, d, px)), vy = fmaf(qy, d, fmaf(r, a, py));
    if (accumulate) {
      gx[base + i] += vx;
      gy[base + i] += vy;
    } else {
      gx[base + i] = vx;
      gy[base + i] = vy;
    }
  }
}

int launch_stat_grad_planar(const float* x, const float* y, const float* coef, int level, int pairs, int c, int hw,
                            float* gx, float* gy, bool accumulate, cudaStream_t st) {
  dim3 grid(ceil_div(hw, 1024), c, pairs);
  stat_grad_planar_kernel<<<grid, 256, 0, st>>>(x, y, coef, level, c, hw, gx, gy, accumulate ? 1 : 0);
  return 1;
}

// ================================================================================================
// NHWC level gradient: statistics term (+ L2pooling backward for levels 1..4), masked by the level's own ReLU.
//   gx_j = [x_j > 0] * ( Px + Qx x_j + R y_j  +  x_j * sum_{o covers j} g(o,j) * gpool_o / pooled_o )
// thread = (input pixel, 16-byte channel vector); POOL = false for the last level.
// ================================================================================================
// SPLIT (bf16x3 carriers): every tensor is a bf16 hi + lo pair; values are hi + lo, outputs are split again.
template <typename T, bool SPLIT>
__device__ inline void load_split(const T* hi, const T* lo, size_t off, float (&v)[Vec<T>::N]) {
  Vec<T>::load(hi + off, v);
  if (SPLIT) {
    float l[Vec<T>::N];
    Vec<T>::load(lo + off, l);
#pragma unroll
    for (int i = 0; i < Vec<T>::N; ++i) v[i] += l[i];
  }
}

template <typename T, bool POOL, bool SPLIT>
__global__ void __launch_bounds__(256)
level_grad_kernel(const T* __restrict__ fx, const T* __restrict__ fy, const T* __restrict__ qpoolx,
                  const T* __restrict__ qpooly, const float* __restrict__ coef, int level, int h, int w, int c,
                  int groups_log2, int ho, int wo, T* __restrict__ gx, T* __restrict__ gy, LevelGradLo<T> lo) {
  // grid = (ceil(w * groups / 256), h, pairs): no divisions in the index math, 32-bit offsets inside an image
  constexpr int V = Vec<T>::N;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int g = idx & ((1 << groups_log2) - 1), ix = idx >> groups_log2;
  if (ix >= w) return;
  const int iy = blockIdx.y, b = blockIdx.z;
  const size_t off = (((size_t)b * h + iy) * w + ix) * c + (size_t)g * V;
  float x[V], y[V], ox[V], oy[V];
  load_split<T, SPLIT>(fx, lo.fx, off, x);
  load_split<T, SPLIT>(fy, lo.fy, off, y);
  float sx[V], sy[V];
  if (POOL) {
    // Pooled outputs whose 3x3 stride-2 window (rows 2o-1..2o+1) contains this pixel: one if the index is even
    // (centre tap, weight 2), two if it is odd (edge taps, weight 1 each).  All four candidate loads are issued
    // unconditionally from clamped addresses and weighted by 0/1 (loads under branches would serialise their latencies).
    const int oy0 = iy >> 1, ox0 = ix >> 1;
    const float my0 = oy0 < ho ? 1.f : 0.f, my1 = ((iy & 1) && oy0 + 1 < ho) ? 1.f : 0.f;
    const float mx0 = ox0 < wo ? 1.f : 0.f, mx1 = ((ix & 1) && ox0 + 1 < wo) ? 1.f : 0.f;
    const int cy0 = min(oy0, ho - 1), cy1 = min(oy0 + 1, ho - 1), cx0 = min(ox0, wo - 1), cx1 = min(ox0 + 1, wo - 1);
    const size_t qbase = (size_t)b * ho * wo * c + (size_t)g * V;
    const int o00 = (cy0 * wo + cx0) * c, o01 = (cy0 * wo + cx1) * c, o10 = (cy1 * wo + cx0) * c, o11 = (cy1 * wo + cx1) * c;
    float a00[V], a01[V], a10[V], a11[V], d00[V], d01[V], d10[V], d11[V];
    load_split<T, SPLIT>(qpoolx, lo.qpoolx, qbase + o00, a00);        // = d loss / d pooled  /  pooled  (kEpiDivide epilogue)
    load_split<T, SPLIT>(qpoolx, lo.qpoolx, qbase + o01, a01);
    load_split<T, SPLIT>(qpoolx, lo.qpoolx, qbase + o10, a10);
    load_split<T, SPLIT>(qpoolx, lo.qpoolx, qbase + o11, a11);
    load_split<T, SPLIT>(qpooly, lo.qpooly, qbase + o00, d00);
    load_split<T, SPLIT>(qpooly, lo.qpooly, qbase + o01, d01);
    load_split<T, SPLIT>(qpooly, lo.qpooly, qbase + o10, d10);
    load_split<T, SPLIT>(qpooly, lo.qpooly, qbase + o11, d11);
    const float m00 = my0 * mx0, m01 = my0 * mx1, m10 = my1 * mx0, m11 = my1 * mx1;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sx[i] = fmaf(m11, a11[i], fmaf(m10, a10[i], fmaf(m01, a01[i], m00 * a00[i])));
      sy[i] = fmaf(m11, d11[i], fmaf(m10, d10[i], fmaf(m01, d01[i], m00 * d00[i])));
    }
  }
  // five coefficient vectors of this (pair, channel group): 16-byte loads from the padded table
  const float* k = coef + (size_t)b * 5 * kCoefStride + coef_offset(level) + g * V;
  float kp[V], kq[V], kr[V], kpy[V], kqy[V];
#pragma unroll
  for (int i = 0; i < V; i += 4) {
    *reinterpret_cast<float4*>(kp + i) = __ldg(reinterpret_cast<const float4*>(k + i));
    *reinterpret_cast<float4*>(kq + i) = __ldg(reinterpret_cast<const float4*>(k + kCoefStride + i));
    *reinterpret_cast<float4*>(kr + i) = __ldg(reinterpret_cast<const float4*>(k + 2 * kCoefStride + i));
    *reinterpret_cast<float4*>(kpy + i) = __ldg(reinterpret_cast<const float4*>(k + 3 * kCoefStride + i));
    *reinterpret_cast<float4*>(kqy + i) = __ldg(reinterpret_cast<const float4*>(k + 4 * kCoefStride + i));
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    ox[i] = fmaf(kq[i], x[i], fmaf(kr[i], y[i], kp[i]));
    oy[i] = fmaf(kqy[i], y[i], fmaf(kr[i], x[i], kpy[i]));
  }
  if (POOL) {
    const float wgt = ((iy & 1) ? 1.f : 2.f) * ((ix & 1) ? 1.f : 2.f) * (1.f / 16.f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      ox[i] = fmaf(wgt * x[i], sx[i], ox[i]);
      oy[i] = fmaf(wgt * y[i], sy[i], oy[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if (!(x[i] > 0.f)) ox[i] = 0.f;
    if (!(y[i] > 0.f)) oy[i] = 0.f;
  }
  Vec<T>::store(gx + off, ox);
  Vec<T>::store(gy + off, oy);
  if (SPLIT) {   // residual carriers: v - round(v)
    float rx[V], ry[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      rx[i] = ox[i] - to_float(from_float<T>(ox[i]));
      ry[i] = oy[i] - to_float(from_float<T>(oy[i]));
    }
    Vec<T>::store(lo.gx + off, rx);
    Vec<T>::store(lo.gy + off, ry);
  }
}

template <typename T>
int launch_level_grad(const T* fx, const T* fy, const T* qpoolx, const T* qpooly, const float* coef, int level, int pairs,
                      int h, int w, int c, T* gx, T* gy, cudaStream_t st, const LevelGradLo<T>* lo) {
  const int groups = c / Vec<T>::N;            // 8 .. 128, a power of two for every VGG level
  int groups_log2 = 0;
  while ((1 << groups_log2) < groups) ++groups_log2;
  if ((1 << groups_log2) != groups || h > 65535 || pairs > 65535) return 0;
  const dim3 grid((unsigned)((w * groups + 255) / 256), (unsigned)h, (unsigned)pairs);
  const LevelGradLo<T> none;
  if (lo) {
    if (qpoolx)
      level_grad_kernel<T, true, true><<<grid, 256, 0, st>>>(fx, fy, qpoolx, qpooly, coef, level, h, w, c, groups_log2,
                                                            pooled(h), pooled(w), gx, gy, *lo);
    else
      level_grad_kernel<T, false, true><<<grid, 256, 0, st>>>(fx, fy, nullptr, nullptr, coef, level, h, w, c, groups_log2, 0,
                                                             0, gx, gy, *lo);
  } else if (qpoolx) {
    level_grad_kernel<T, true, false><<<grid, 256, 0, st>>>(fx, fy, qpoolx, qpooly, coef, level, h, w, c, groups_log2,
                                                           pooled(h), pooled(w), gx, gy, none);
  } else {
    level_grad_kernel<T, false, false><<<grid, 256, 0, st>>>(fx, fy, nullptr, nullptr, coef, level, h, w, c, groups_log2, 0,
                                                            0, gx, gy, none);
  }
  return 1;
}
template int launch_level_grad<float>(const float*, const float*, const float*, const float*, const float*, int, int, int, int,
                                      int, float*, float*, cudaStream_t, const LevelGradLo<float>*);
template int launch_level_grad<bf16>(const bf16*, const bf16*, const bf16*, const bf16*, const float*, int, int, int, int, int,
                                     bf16*, bf16*, cudaStream_t, const LevelGradLo<bf16>*);

// ================================================================================================
// conv1_1 backward + input normalisation: gz [n,h,w,64] (gradient w.r.t. the pre-ReLU output of conv1_1, already
// masked) -> gradient w.r.t. the RAW frames, NCHW fp32 (images [0,nx) -> gx, [nx,n) -> gy).
// Weights [9][64][3] arrive as a kernel parameter (constant bank).
// ================================================================================================
// Thread = one input pixel.  ncu on B200 (512x512, 32 images, bf16): weights as scalar shared-memory reads 2.9 ms per
// launch, this version (weights as compile-time offsets into the constant bank) 2.1 ms.  Two pixels per thread with
// float4 weights from shared memory, and staging the gradient rows in shared memory, were tried and did not move the
// step time; the kernel is far from the FP32 roofline (0.4 ms) -- a tensor-core N=16 variant is the next step.
template <typename T>
__global__ void __launch_bounds__(128)
conv1_1_bwd_kernel(const T* __restrict__ gz, const T* __restrict__ gz_lo, const __grid_constant__ Conv11BwdWeights wb,
                   int nx, int h, int w, float* __restrict__ gx, float* __restrict__ gy) {
  constexpr int V = Vec<T>::N;
  const int img = blockIdx.z, y = blockIdx.y, x = blockIdx.x * 128 + threadIdx.x;
  if (x >= w) return;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int oy = y + 1 - r;
    if (oy < 0 || oy >= h) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ox = x + 1 - s;
      if (ox < 0 || ox >= w) continue;
      const size_t soff = (((size_t)img * h + oy) * w + ox) * 64;
      const T* src = gz + soff;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += V) {
        float v[V];
        Vec<T>::load(src + c0, v);
        if (gz_lo) {   // bf16x3 carrier
          float l[V];
          Vec<T>::load(gz_lo + soff + c0, l);
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] += l[i];
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {     // weights: compile-time offsets into the constant bank
          acc[0] = fmaf(v[i], wb.w[r * 3 + s][c0 + i][0], acc[0]);
          acc[1] = fmaf(v[i], wb.w[r * 3 + s][c0 + i][1], acc[1]);
          acc[2] = fmaf(v[i], wb.w[r * 3 + s][c0 + i][2], acc[2]);
        }
      }
    }
  }
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  float* dst = img < nx ? gx + (size_t)img * 3 * h * w : gy + (size_t)(img - nx) * 3 * h * w;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) dst[((size_t)ci * h + y) * w + x] = acc[ci] / stdv[ci];
}

template <typename T>
int launch_conv1_1_bwd(const T* gz, const Conv11BwdWeights& wb, int nx, int n, int h, int w, float* gx, float* gy,
                       cudaStream_t st, const T* gz_lo) {
  dim3 grid(ceil_div(w, 128), h, n);
  conv1_1_bwd_kernel<T><<<grid, 128, 0, st>>>(gz, gz_lo, wb, nx, h, w, gx, gy);
  return 1;
}
template int launch_conv1_1_bwd<float>(const float*, const Conv11BwdWeights&, int, int, int, int, float*, float*, cudaStream_t,
                                       const float*);
template int launch_conv1_1_bwd<bf16>(const bf16*, const Conv11BwdWeights&, int, int, int, int, float*, float*, cudaStream_t,
                                      const bf16*);

// ================================================================================================
// weight packing for the data-gradient convolutions
// ================================================================================================
// w OIHW [cout][cin][3][3] ->  w_f32 [tap'][cout][cin],  w_hi [cin][tap'][cout] (K-major rows for the tcgen05 B
// operand), tap' = 8 - tap (180-degree rotation), roles of cin / cout swapped.
__global__ void pack_w_bwd_kernel(const float* __restrict__ w, int cin, int cout, float* __restrict__ w_f32,
                                  bf16* __restrict__ w_hi, bf16* __restrict__ w_lo) {
  const size_t total = (size_t)cout * cin * 9;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9);
    const int ci = (int)((i / 9) % cin);
    const int co = (int)(i / (9 * (size_t)cin));
    const float v = w[i];
    const int tb = 8 - tap;
    if (w_f32) w_f32[((size_t)tb * cout + co) * cin + ci] = v;
    const bf16 hi = __float2bfloat16_rn(v);
    w_hi[(size_t)ci * 9 * cout + (size_t)tb * cout + co] = hi;
    if (w_lo) w_lo[(size_t)ci * 9 * cout + (size_t)tb * cout + co] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}
// conv1_1: w [64][3][3][3] -> [tap][co][ci] (indexed by the FORWARD tap: input pixel p feeds output p - shift(tap))
__global__ void pack_w11_bwd_kernel(const float* __restrict__ w, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 27) return;
  const int tap = i % 9, ci = (i / 9) % 3, co = i / 27;
  out[(tap * 64 + co) * 3 + ci] = w[i];
}

int launch_pack_w_bwd(const float* w_oihw, int cin, int cout, float* w_f32, bf16* w_hi, cudaStream_t st, bf16* w_lo) {
  pack_w_bwd_kernel<<<1024, 256, 0, st>>>(w_oihw, cin, cout, w_f32, w_hi, w_lo);
  return 1;
}
int launch_pack_w11_bwd(const float* w_oihw, float* out, cudaStream_t st) {
  pack_w11_bwd_kernel<<<(64 * 27 + 255) / 256, 256, 0, st>>>(w_oihw, out);
  return 1;
}

}  // namespace nqa
