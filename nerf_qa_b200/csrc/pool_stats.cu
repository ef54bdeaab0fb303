// L2pooling, DISTS statistics + finalize, layout transforms.  All HBM-bound: 16-byte vector accesses,
// coalesced along the NHWC channel axis, fp32 accumulation, fp64 only in the tiny finalize step.
//
// Reference semantics: nerf_qa/DISTS_pytorch/DISTS_pt.py:11-25 (L2pooling), :113-148 (statistics).
#include "common.cuh"
#include "kernels.h"

namespace nqa {

// ================================================================================================
// DISTS statistics, NHWC.  grid = (chunks, pairs); each block owns a contiguous pixel range of one
// frame pair and all C channels; thread = (pixel lane, channel vector).  Per thread at most a few
// hundred sequential fp32 adds per accumulator (SURVEY.md §7 "variance numerics").
// ================================================================================================
constexpr int kStatsThreads = 256;

template <typename T, bool SPLIT>
__global__ void __launch_bounds__(kStatsThreads)
stats_nhwc_kernel(const T* __restrict__ fx, const T* __restrict__ fy, const T* __restrict__ fx_lo,
                  const T* __restrict__ fy_lo, int hw, int c, int chunks, float* __restrict__ partials) {
  constexpr int V = Vec<T>::N;
  const int groups = c / V;                 // channel vectors per pixel (<= 256)
  const int ppi = kStatsThreads / groups;   // pixels per block iteration
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int per = ceil_div(hw, chunks);
  const int p0 = chunk * per, p1 = min(hw, p0 + per);

  float sx[V], sy[V], sxx[V], syy[V], sxy[V];
#pragma unroll
  for (int i = 0; i < V; ++i) sx[i] = sy[i] = sxx[i] = syy[i] = sxy[i] = 0.f;

  if (pl < ppi) {
    const size_t off = (size_t)b * hw * c + (size_t)g * V;
    const T* px = fx + off;
    const T* py = fy + off;
    // pivots = the channel's first pixel (kernels.h: shifted one-pass sums)
    float pvx[V], pvy[V];
    Vec<T>::load(px, pvx);
    Vec<T>::load(py, pvy);
    if (SPLIT) {
      float lx[V], ly[V];
      Vec<T>::load(fx_lo + off, lx);
      Vec<T>::load(fy_lo + off, ly);
#pragma unroll
      for (int i = 0; i < V; ++i) { pvx[i] += lx[i]; pvy[i] += ly[i]; }
    }
    if (chunk == 0 && pl == 0) {
      float* piv = partials + (size_t)gridDim.y * chunks * 5 * c + (size_t)b * 2 * c + (size_t)g * V;
#pragma unroll
      for (int i = 0; i < V; ++i) { piv[i] = pvx[i]; piv[c + i] = pvy[i]; }
    }
#pragma unroll 4
    for (int p = p0 + pl; p < p1; p += ppi) {
      float vx[V], vy[V];
      Vec<T>::load_stream(px + (size_t)p * c, vx);
      Vec<T>::load_stream(py + (size_t)p * c, vy);
      if (SPLIT) {  // bf16x3 carrier: value = hi + lo
        float lx[V], ly[V];
        Vec<T>::load_stream(fx_lo + off + (size_t)p * c, lx);
        Vec<T>::load_stream(fy_lo + off + (size_t)p * c, ly);
#pragma unroll
        for (int i = 0; i < V; ++i) { vx[i] += lx[i]; vy[i] += ly[i]; }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) { vx[i] -= pvx[i]; vy[i] -= pvy[i]; }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        sx[i] += vx[i];
        sy[i] += vy[i];
        sxx[i] = fmaf(vx[i], vx[i], sxx[i]);
        syy[i] = fmaf(vy[i], vy[i], syy[i]);
        sxy[i] = fmaf(vx[i], vy[i], sxy[i]);
      }
    }
  }

  __shared__ float red[kStatsThreads * V];
  float* outp = partials + ((size_t)b * chunks + chunk) * 5 * c;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const float* src = s == 0 ? sx : s == 1 ? sy : s == 2 ? sxx : s == 3 ? syy : sxy;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) red[threadIdx.x * V + i] = src[i];
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += kStatsThreads) {
      const int gg = ch / V, ii = ch % V;
      float acc = 0.f;
      for (int q = 0; q < ppi; ++q) acc += red[(q * groups + gg) * V + ii];
      outp[s * c + ch] = acc;
    }
  }
}

// NOTE: the partition depends on the level shape only, never on the batch, so that a frame's score is
// bit-identical whatever micro-batch it travels in.
int stats_chunks_nhwc(int hw, int c, int elem_bytes, int /*pairs*/) {
  const int pairs = 1;
  const int v = 16 / elem_bytes;
  const int groups = c / v;
  const int ppi = kStatsThreads / groups;
  // <= 256 sequential adds per accumulator, but at least ~2 waves of blocks when the level allows it
  int chunks = ceil_div(hw, ppi * 256);
  const int want = 64 / pairs;  // enough blocks per frame pair to spread a small level over the SMs
  const int cap = ceil_div(hw, ppi * 8);
  if (chunks < want) chunks = want < cap ? want : cap;
  if (chunks < 1) chunks = 1;
  return chunks;
}

template <typename T>
int launch_stats_nhwc(const T* fx, const T* fy, const T* fx_lo, const T* fy_lo, int pairs, int hw, int c, int chunks,
                      float* partials, cudaStream_t st) {
  dim3 grid(chunks, pairs);
  if (fx_lo)
    stats_nhwc_kernel<T, true><<<grid, kStatsThreads, 0, st>>>(fx, fy, fx_lo, fy_lo, hw, c, chunks, partials);
  else
    stats_nhwc_kernel<T, false><<<grid, kStatsThreads, 0, st>>>(fx, fy, nullptr, nullptr, hw, c, chunks, partials);
  return 1;
}
template int launch_stats_nhwc<float>(const float*, const float*, const float*, const float*, int, int, int, int,
                                      float*, cudaStream_t);
template int launch_stats_nhwc<bf16>(const bf16*, const bf16*, const bf16*, const bf16*, int, int, int, int, float*,
                                     cudaStream_t);

// ================================================================================================
// DISTS statistics on NCHW fp32 planes (level 0 = raw frames; every level of forward_from_feats).
// grid = (chunks, C, pairs).
// ================================================================================================
__global__ void __launch_bounds__(256)
stats_planar_kernel(const float* __restrict__ x, const float* __restrict__ y, int c, int hw, int chunks,
                    float* __restrict__ partials) {
  const int chunk = blockIdx.x, ch = blockIdx.y, b = blockIdx.z;
  const int per = ceil_div(ceil_div(hw, chunks), 4) * 4;  // keep 16-byte alignment of chunk starts when hw % 4 == 0
  const int p0 = chunk * per, p1 = min(hw, p0 + per);
  const float* px = x + ((size_t)b * c + ch) * hw;
  const float* py = y + ((size_t)b * c + ch) * hw;
  const float pvx = __ldg(px), pvy = __ldg(py);   // pivots = the channel's first pixel (kernels.h: shifted one-pass sums)
  if (chunk == 0 && threadIdx.x == 0) {
    float* piv = partials + (size_t)gridDim.z * chunks * 5 * c + (size_t)b * 2 * c;
    piv[ch] = pvx;
    piv[c + ch] = pvy;
  }
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const bool vec_ok = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(px) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(py) & 15) == 0);
  if (vec_ok) {
    const int n4 = (p1 - p0) / 4;
    const float4* qx = reinterpret_cast<const float4*>(px + p0);
    const float4* qy = reinterpret_cast<const float4*>(py + p0);
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 a = __ldcs(qx + i), d = __ldcs(qy + i);
      a.x -= pvx; a.y -= pvx; a.z -= pvx; a.w -= pvx;
      d.x -= pvy; d.y -= pvy; d.z -= pvy; d.w -= pvy;
      s[0] += (a.x + a.y) + (a.z + a.w);
      s[1] += (d.x + d.y) + (d.z + d.w);
      s[2] += fmaf(a.x, a.x, a.y * a.y) + fmaf(a.z, a.z, a.w * a.w);
      s[3] += fmaf(d.x, d.x, d.y * d.y) + fmaf(d.z, d.z, d.w * d.w);
      s[4] += fmaf(a.x, d.x, a.y * d.y) + fmaf(a.z, d.z, a.w * d.w);
    }
    for (int p = p0 + n4 * 4 + threadIdx.x; p < p1; p += 256) {
      const float a = px[p] - pvx, d = py[p] - pvy;
      s[0] += a; s[1] += d; s[2] = fmaf(a, a, s[2]); s[3] = fmaf(d, d, s[3]); s[4] = fmaf(a, d, s[4]);
    }
  } else {
    for (int p = p0 + threadIdx.x; p < p1; p += 256) {
      const float a = px[p] - pvx, d = py[p] - pvy;
      s[0] += a; s[1] += d; s[2] = fmaf(a, a, s[2]); s[3] = fmaf(d, d, s[3]); s[4] = fmaf(a, d, s[4]);
    }
  }
  __shared__ float sm[5][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const float v = warp_sum(s[i]);
    if (lane == 0) sm[i][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += sm[threadIdx.x][q];
    partials[(((size_t)b * chunks + chunk) * 5 + threadIdx.x) * c + ch] = v;
  }
}

int stats_chunks_planar(int hw, int c, int /*pairs*/) {
  const int pairs = 1;
  // 256 threads x float4 x <=64 iterations per accumulator
  int chunks = ceil_div(hw, 256 * 4 * 64);
  const int want = ceil_div(64, pairs * c);
  const int cap = ceil_div(hw, 256 * 4);
  if (chunks < want) chunks = want < cap ? want : cap;
  if (chunks < 1) chunks = 1;
  return chunks;
}

int launch_stats_planar(const float* x, const float* y, int pairs, int c, int hw, int chunks, float* partials,
                        cudaStream_t st) {
  dim3 grid(chunks, c, pairs);
  stats_planar_kernel<<<grid, 256, 0, st>>>(x, y, c, hw, chunks, partials);
  return 1;
}

// ================================================================================================
// Finalize: moments -> S1,S2 -> alpha/beta-weighted sum -> score (DISTS_pt.py:127-148).
// Stage 1, grid = (ceil(1475/64), pairs), 1024 threads = 64 channels x 16 chunk lanes: chunk partials are
// combined in fp64 in a fixed order (lane q takes chunks q, q+16, ...; lanes combined 0..15 in sequence), so
// the result is deterministic and batch-independent.  Stage 2, grid = pairs: ordered sum of the 1475
// per-channel terms.
// ================================================================================================
constexpr int kFinCh = 64;
constexpr int kFinLanes = 16;

__global__ void __launch_bounds__(kFinCh * kFinLanes)
dists_terms_kernel(FinalizeParams p, const float* __restrict__ alpha, const float* __restrict__ beta,
                   double* __restrict__ terms, float* __restrict__ s1s2, float* __restrict__ moments) {
  __shared__ double sm[kFinLanes][kFinCh][5];
  const int b = blockIdx.y;
  const int cl = threadIdx.x % kFinCh, q = threadIdx.x / kFinCh;
  const int c = blockIdx.x * kFinCh + cl;
  double s[5] = {0, 0, 0, 0, 0};
  LevelPartials lv = p.lv[0];
  if (c < kTotalChannels) {
    const int k = c < 3 ? 0 : c < 67 ? 1 : c < 195 ? 2 : c < 451 ? 3 : c < 963 ? 4 : 5;
    lv = p.lv[k];
    const int cc = c - lv.choff;
    const float* base = lv.partials + (size_t)b * lv.chunks * 5 * lv.C + cc;
    for (int ch = q; ch < lv.chunks; ch += kFinLanes) {
#pragma unroll
      for (int i = 0; i < 5; ++i) s[i] += (double)base[((size_t)ch * 5 + i) * lv.C];
    }
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) sm[q][cl][i] = s[i];
  __syncthreads();
  if (q == 0 && c < kTotalChannels) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double t = sm[0][cl][i];
#pragma unroll
      for (int l = 1; l < kFinLanes; ++l) t += sm[l][cl][i];
      s[i] = t;
    }
    // sums are of the pivot-shifted values: variances / covariance are shift invariant, the mean adds the pivot back
    const double dx = s[0] * lv.inv_n, dy = s[1] * lv.inv_n;
    const double vx = s[2] * lv.inv_n - dx * dx, vy = s[3] * lv.inv_n - dy * dy;
    const double cov = s[4] * lv.inv_n - dx * dy;
    const float* piv = lv.piv + (size_t)b * 2 * lv.C + (c - lv.choff);
    const double mx = (double)piv[0] + dx, my = (double)piv[lv.C] + dy;
    const double c1 = 1e-6, c2 = 1e-6;
    const double s1 = (2.0 * mx * my + c1) / (mx * mx + my * my + c1);
    const double s2 = (2.0 * cov + c2) / (vx + vy + c2);
    if (s1s2) {
      s1s2[((size_t)b * 2 + 0) * kTotalChannels + c] = (float)s1;
      s1s2[((size_t)b * 2 + 1) * kTotalChannels + c] = (float)s2;
    }
    if (moments) {
      float* m = moments + ((size_t)b * kTotalChannels + c) * 5;
      m[0] = (float)mx; m[1] = (float)my; m[2] = (float)vx; m[3] = (float)vy; m[4] = (float)cov;
    }
    terms[(size_t)b * kTotalChannels + c] = (double)alpha[c] * s1 + (double)beta[c] * s2;
  }
}

__global__ void __launch_bounds__(256)
dists_score_kernel(const double* __restrict__ terms, const float* __restrict__ alpha, const float* __restrict__ beta,
                   int normalize, float* __restrict__ score) {
  __shared__ double smd[33];
  const int b = blockIdx.x;
  double wpart = 0.0, tpart = 0.0;
  for (int c = threadIdx.x; c < kTotalChannels; c += 256) {
    wpart += (double)alpha[c] + (double)beta[c];
    tpart += terms[(size_t)b * kTotalChannels + c];
  }
  const double wsum = block_sum<double>(wpart, smd);
  const double total = block_sum<double>(tpart, smd);
  if (threadIdx.x == 0) score[b] = (float)(1.0 - (normalize ? total / wsum : total));
}

int launch_dists_finalize(const FinalizeParams& p, const float* alpha, const float* beta, bool normalize, int pairs,
                          float* score, float* s1s2, float* moments, double* terms, cudaStream_t st) {
  dim3 grid(ceil_div(kTotalChannels, kFinCh), pairs);
  dists_terms_kernel<<<grid, kFinCh * kFinLanes, 0, st>>>(p, alpha, beta, terms, s1s2, moments);
  dists_score_kernel<<<pairs, 256, 0, st>>>(terms, alpha, beta, normalize ? 1 : 0, score);
  return 2;
}

// ================================================================================================
// L2pooling (DISTS_pt.py:22-25): out = sqrt(sum_{3x3} g * x^2 + 1e-12), g = [1 2 1]^T[1 2 1]/16,
// stride 2, zero pad 1.  NHWC: thread = (output pixel, 16-byte channel vector).
// ================================================================================================
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(256)
l2pool_nhwc_kernel(const T* __restrict__ in, const T* __restrict__ in_lo, T* __restrict__ out, T* __restrict__ out_lo,
                   int h, int w, int c, int groups, int ho, int wo) {
  // grid = (ceil(wo * groups / 256), ho, n): one division per thread, none in 64 bits.  The nine taps are loaded
  // unconditionally from clamped addresses with weight 0 outside the image (loads under branches serialise).
  constexpr int V = Vec<T>::N;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int ox = idx / groups, g = idx - ox * groups;
  if (ox >= wo) return;
  const int oy = blockIdx.y, img = blockIdx.z;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  const size_t boff = (size_t)img * h * w * c + (size_t)g * V;
  const T* base = in + boff;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = 2 * oy - 1 + r;
    const bool rowok = iy >= 0 && iy < h;
    const int cy = min(max(iy, 0), h - 1);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ix = 2 * ox - 1 + s;
      const bool ok = rowok && ix >= 0 && ix < w;
      const int cx = min(max(ix, 0), w - 1);
      const float wt = ok ? ((r == 1) ? 2.f : 1.f) * ((s == 1) ? 2.f : 1.f) * (1.f / 16.f) : 0.f;
      float v[V];
      Vec<T>::load(base + ((size_t)cy * w + cx) * c, v);
      if (SPLIT) {
        float l[V];
        Vec<T>::load(in_lo + boff + ((size_t)cy * w + cx) * c, l);
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] += l[i];
      }
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = fmaf(wt, v[i] * v[i], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = sqrtf(acc[i] + 1e-12f);
  const size_t ooff = (((size_t)img * ho + oy) * wo + ox) * c + (size_t)g * V;
  Vec<T>::store(out + ooff, acc);
  if (SPLIT) {
    float l[V];
#pragma unroll
    for (int i = 0; i < V; ++i) l[i] = acc[i] - to_float(from_float<T>(acc[i]));
    Vec<T>::store(out_lo + ooff, l);
  }
}

template <typename T>
int launch_l2pool_nhwc(const T* in, const T* in_lo, T* out, T* out_lo, int n, int h, int w, int c, cudaStream_t st) {
  const int ho = pooled(h), wo = pooled(w), groups = c / Vec<T>::N;
  if (ho > 65535 || n > 65535) return 0;
  const dim3 grid((unsigned)((wo * groups + 255) / 256), (unsigned)ho, (unsigned)n);
  if (in_lo)
    l2pool_nhwc_kernel<T, true><<<grid, 256, 0, st>>>(in, in_lo, out, out_lo, h, w, c, groups, ho, wo);
  else
    l2pool_nhwc_kernel<T, false><<<grid, 256, 0, st>>>(in, nullptr, out, nullptr, h, w, c, groups, ho, wo);
  return 1;
}
template int launch_l2pool_nhwc<float>(const float*, const float*, float*, float*, int, int, int, int, cudaStream_t);
template int launch_l2pool_nhwc<bf16>(const bf16*, const bf16*, bf16*, bf16*, int, int, int, int, cudaStream_t);

// ================================================================================================
// Fused L2pooling + DISTS statistics of one pyramid level (levels 1..4 of the scoring path): the level's
// features of x and y are read ONCE; the kernel emits the pooled input of the next stage for both images
// and the five moment sums per channel.  grid = (col blocks, row strips, pairs), 256 threads =
// (pooled column lane, 16-byte channel vector).  A thread walks down R pooled rows of its pooled column:
// per input row it loads the three columns 2ox-1..2ox+1 of x and y, forms the horizontal [1 2 1] sum of
// squares, and combines rows with a rolling vertical [1 2 1] (the odd row is shared by two pooled rows).
// Statistics are accumulated on the 2x2 pixels a pooled pixel owns ((2oy, 2oy+1) x (2ox, 2ox+1)), so every
// input pixel is counted exactly once; <= 4R sequential fp32 adds per accumulator.
// ================================================================================================
// 4 channels per thread (8-byte bf16 / 16-byte fp32 vectors) keep the kernel at ~80 registers (three resident
// blocks per SM).  The kernel is as much issue-bound as HBM-bound (~25 instructions per x/y element pair), so the
// fp32 math uses the sm_100 packed f32x2 instructions (two channels per FFMA2/FADD2/FMUL2).
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  __device__ static inline void load(const float* p, float2 (&v)[2]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = make_float2(t.x, t.y); v[1] = make_float2(t.z, t.w);
  }
  __device__ static inline void store(float* p, const float2 (&v)[2]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
  }
  __device__ static inline float root(float v) { return sqrtf(v); }
};
template <> struct Vec4<bf16> {
  __device__ static inline void load(const bf16* p, float2 (&v)[2]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = make_float2(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u));
    v[1] = make_float2(__uint_as_float(t.y << 16), __uint_as_float(t.y & 0xffff0000u));
  }
  __device__ static inline void store(bf16* p, const float2 (&v)[2]) {
    uint2 t;
    t.x = Vec<bf16>::pack2(v[0].x, v[0].y);
    t.y = Vec<bf16>::pack2(v[1].x, v[1].y);
    *reinterpret_cast<uint2*>(p) = t;
  }
  // the result is rounded to bf16 (8 bits of mantissa): the 1-ulp-fp32 MUFU square root is exact enough
  __device__ static inline float root(float v) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
  }
};

template <typename T>
__global__ void __launch_bounds__(256, 3)
pool_stats_kernel(const T* __restrict__ fx, const T* __restrict__ fy, T* __restrict__ px_out, T* __restrict__ py_out,
                  int h, int w, int c, int ho, int wo, int R, int chunks, float* __restrict__ partials) {
  constexpr int V = 4;
  const int G = c / V, PL = 256 / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int b = blockIdx.z;
  const int ox = blockIdx.x * PL + pl;
  const int oy0 = blockIdx.y * R, oy1 = min(ho, oy0 + R);
  const bool active = ox < wo;
  const float2 zero2 = make_float2(0.f, 0.f), two2 = make_float2(2.f, 2.f);

  float2 s[5][2];
#pragma unroll
  for (int q = 0; q < 5; ++q) s[q][0] = s[q][1] = zero2;

  if (active) {
    const size_t ioff = (size_t)b * h * w * c + (size_t)g * V;
    const size_t ooff = (size_t)b * ho * wo * c + (size_t)g * V;
    const T* bx = fx + ioff;
    const T* by = fy + ioff;
    const bool c0 = 2 * ox - 1 >= 0, c2 = 2 * ox + 1 < w;   // column 2ox always exists
    // pivots = the channel's first pixel, negated (kernels.h: shifted one-pass sums)
    float2 npx[2], npy[2];
    Vec4<T>::load(bx, npx);
    Vec4<T>::load(by, npy);
    if (blockIdx.x == 0 && blockIdx.y == 0 && pl == 0) {
      float* piv = partials + (size_t)gridDim.z * chunks * 5 * c + (size_t)b * 2 * c + (size_t)g * V;
      piv[0] = npx[0].x; piv[1] = npx[0].y; piv[2] = npx[1].x; piv[3] = npx[1].y;
      piv[c + 0] = npy[0].x; piv[c + 1] = npy[0].y; piv[c + 2] = npy[1].x; piv[c + 3] = npy[1].y;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      npx[i] = make_float2(-npx[i].x, -npx[i].y);
      npy[i] = make_float2(-npy[i].x, -npy[i].y);
    }

    // horizontal [1 2 1] of squares of input row iy for both images (+ statistics on the owned columns)
    // rx / ry point at column 2ox of the input row; rows outside the image contribute zeros
    auto row = [&](const T* rx, const T* ry, bool in_image, bool own, float2 (&hx)[2], float2 (&hy)[2]) {
      if (!in_image) {
        hx[0] = hx[1] = hy[0] = hy[1] = zero2;
        return;
      }
      float2 x0[2] = {zero2, zero2}, x1[2], x2[2] = {zero2, zero2};
      float2 y0[2] = {zero2, zero2}, y1[2], y2[2] = {zero2, zero2};
      Vec4<T>::load(rx, x1);
      Vec4<T>::load(ry, y1);
      if (c0) { Vec4<T>::load(rx - c, x0); Vec4<T>::load(ry - c, y0); }
      if (c2) { Vec4<T>::load(rx + c, x2); Vec4<T>::load(ry + c, y2); }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float2 xx1 = __fmul2_rn(x1[i], x1[i]), yy1 = __fmul2_rn(y1[i], y1[i]);
        const float2 xx2 = __fmul2_rn(x2[i], x2[i]), yy2 = __fmul2_rn(y2[i], y2[i]);
        hx[i] = __ffma2_rn(x0[i], x0[i], __ffma2_rn(two2, xx1, xx2));
        hy[i] = __ffma2_rn(y0[i], y0[i], __ffma2_rn(two2, yy1, yy2));
        if (own) {
          // column 2ox+1 outside the image must not be counted: its shifted value is forced to zero
          const float2 dx1 = __fadd2_rn(x1[i], npx[i]), dy1 = __fadd2_rn(y1[i], npy[i]);
          const float2 dx2 = c2 ? __fadd2_rn(x2[i], npx[i]) : zero2, dy2 = c2 ? __fadd2_rn(y2[i], npy[i]) : zero2;
          s[0][i] = __fadd2_rn(s[0][i], __fadd2_rn(dx1, dx2));
          s[1][i] = __fadd2_rn(s[1][i], __fadd2_rn(dy1, dy2));
          s[2][i] = __ffma2_rn(dx2, dx2, __ffma2_rn(dx1, dx1, s[2][i]));
          s[3][i] = __ffma2_rn(dy2, dy2, __ffma2_rn(dy1, dy1, s[3][i]));
          s[4][i] = __ffma2_rn(dx2, dy2, __ffma2_rn(dx1, dy1, s[4][i]));
        }
      }
    };

    // running row pointers (column 2ox of row 2oy - 1); one 64-bit add per input row instead of a multiply
    const ptrdiff_t rstride = (ptrdiff_t)w * c;
    const T* rx = bx + ((ptrdiff_t)(2 * oy0 - 1) * w + 2 * ox) * c;
    const T* ry = by + ((ptrdiff_t)(2 * oy0 - 1) * w + 2 * ox) * c;
    T* ox_ptr = px_out + ooff + ((size_t)oy0 * wo + ox) * c;
    T* oy_ptr = py_out + ooff + ((size_t)oy0 * wo + ox) * c;
    const ptrdiff_t ostride = (ptrdiff_t)wo * c;
    float2 px[2], py[2];   // horizontal sums of the odd row above (2oy - 1)
    row(rx, ry, oy0 > 0, false, px, py);
    const float2 sixteenth = make_float2(1.f / 16.f, 1.f / 16.f), eps = make_float2(1e-12f, 1e-12f);
#pragma unroll 1
    for (int oy = oy0; oy < oy1; ++oy) {
      float2 ax[2], ay[2], bx2[2], by2[2];
      row(rx + rstride, ry + rstride, true, true, ax, ay);               // row 2oy always exists
      rx += 2 * rstride;
      ry += 2 * rstride;
      row(rx, ry, 2 * oy + 1 < h, true, bx2, by2);
      float2 ovx[2], ovy[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float2 tx = __ffma2_rn(__fadd2_rn(__ffma2_rn(two2, ax[i], px[i]), bx2[i]), sixteenth, eps);
        const float2 ty = __ffma2_rn(__fadd2_rn(__ffma2_rn(two2, ay[i], py[i]), by2[i]), sixteenth, eps);
        ovx[i] = make_float2(Vec4<T>::root(tx.x), Vec4<T>::root(tx.y));
        ovy[i] = make_float2(Vec4<T>::root(ty.x), Vec4<T>::root(ty.y));
        px[i] = bx2[i];
        py[i] = by2[i];
      }
      Vec4<T>::store(ox_ptr, ovx);
      Vec4<T>::store(oy_ptr, ovy);
      ox_ptr += ostride;
      oy_ptr += ostride;
    }
  }

  __shared__ float red[256 * V];
  const int chunk = blockIdx.y * gridDim.x + blockIdx.x;
  float* outp = partials + ((size_t)b * chunks + chunk) * 5 * c;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    __syncthreads();
    *reinterpret_cast<float4*>(&red[threadIdx.x * V]) = make_float4(s[q][0].x, s[q][0].y, s[q][1].x, s[q][1].y);
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += 256) {
      const int gg = ch / V, ii = ch % V;
      float acc = 0.f;
      for (int t = 0; t < PL; ++t) acc += red[(t * G + gg) * V + ii];
      outp[q * c + ch] = acc;
    }
  }
}

// Partition of one level for pool_stats_kernel; depends on the level shape only (batch independence).
void pool_stats_plan(int h, int w, int c, int elem_bytes, int* rows_per_block, int* col_blocks, int* strips) {
  (void)elem_bytes;
  const int groups = c / 4, pl = 256 / groups;   // 4 channels per thread
  const int ho = pooled(h), wo = pooled(w);
  const int cb = ceil_div(wo, pl);
  int r = 16;
  while (r > 2 && cb * ceil_div(ho, r) < 256) r >>= 1;
  *rows_per_block = r;
  *col_blocks = cb;
  *strips = ceil_div(ho, r);
}

int pool_stats_chunks(int h, int w, int c, int elem_bytes) {
  int r, cb, st;
  pool_stats_plan(h, w, c, elem_bytes, &r, &cb, &st);
  return cb * st;
}

template <typename T>
int launch_pool_stats(const T* fx, const T* fy, T* px_out, T* py_out, int pairs, int h, int w, int c, int chunks,
                      float* partials, cudaStream_t st, bool allow_tma) {
  if constexpr (std::is_same<T, bf16>::value) {
    if (allow_tma) {   // NQA_POOL_STATS_TMA (per context, default on); off selects the register-fed kernel
      const int rc = launch_pool_stats_tma(fx, fy, px_out, py_out, pairs, h, w, c, chunks, partials, st);
      if (rc > 0) return rc;
    }
  }
  int r, cb, strips;
  pool_stats_plan(h, w, c, (int)sizeof(T), &r, &cb, &strips);
  dim3 grid(cb, strips, pairs);
  pool_stats_kernel<T><<<grid, 256, 0, st>>>(fx, fy, px_out, py_out, h, w, c, pooled(h), pooled(w), r, chunks, partials);
  return 1;
}
template int launch_pool_stats<float>(const float*, const float*, float*, float*, int, int, int, int, int, float*, cudaStream_t, bool);
template int launch_pool_stats<bf16>(const bf16*, const bf16*, bf16*, bf16*, int, int, int, int, int, float*, cudaStream_t, bool);

// Standalone NCHW fp32 L2pooling (the nn.Module surface): thread = one output element.
__global__ void __launch_bounds__(256)
l2pool_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int h, int w, int ho, int wo) {
  const size_t total = (size_t)planes * ho * wo;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % wo);
    size_t t = idx / wo;
    const int oy = (int)(t % ho);
    const size_t pl = t / ho;
    const float* base = in + pl * h * w;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = 2 * oy - 1 + r;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = 2 * ox - 1 + s;
        if (ix < 0 || ix >= w) continue;
        const float wt = ((r == 1) ? 2.f : 1.f) * ((s == 1) ? 2.f : 1.f) * (1.f / 16.f);
        const float v = __ldg(base + (size_t)iy * w + ix);
        acc = fmaf(wt, v * v, acc);
      }
    }
    out[idx] = sqrtf(acc + 1e-12f);
  }
}

int launch_l2pool_nchw(const float* in, float* out, int n, int c, int h, int w, cudaStream_t st) {
  const int ho = pooled(h), wo = pooled(w);
  const size_t total = (size_t)n * c * ho * wo;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  l2pool_nchw_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, n * c, h, w, ho, wo);
  return 1;
}

// ================================================================================================
// NHWC (internal) -> NCHW fp32 (API surface of forward_once), tiled transpose through smem.
// grid = (ceil(hw/32), ceil(c/32), n), block = (32, 8).
// ================================================================================================
template <typename T, bool SPLIT, typename TO>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T* __restrict__ in, const T* __restrict__ in_lo, TO* __restrict__ out, int c, int hw) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const T* src = in + (size_t)img * hw * c;
  const T* src_lo = SPLIT ? in_lo + (size_t)img * hw * c : nullptr;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, ch = c0 + threadIdx.x;
    float v = 0.f;
    if (p < hw && ch < c) {
      v = to_float(src[(size_t)p * c + ch]);
      if (SPLIT) v += to_float(src_lo[(size_t)p * c + ch]);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  TO* dst = out + (size_t)img * c * hw;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ch = c0 + i, p = p0 + threadIdx.x;
    if (p < hw && ch < c) dst[(size_t)ch * hw + p] = from_float<TO>(tile[threadIdx.x][i]);
  }
}

// NCHW fp32 -> NHWC T (optionally hi/lo split); op-level tests only, not on the scoring path.
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, T* __restrict__ out_lo, int c, int hw) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + (size_t)img * c * hw;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ch = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && ch < c) ? src[(size_t)ch * hw + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, ch = c0 + threadIdx.x;
    if (p < hw && ch < c) {
      const float v = tile[threadIdx.x][i];
      const T hi = from_float<T>(v);
      out[((size_t)img * hw + p) * c + ch] = hi;
      if (out_lo) out_lo[((size_t)img * hw + p) * c + ch] = from_float<T>(v - to_float(hi));
    }
  }
}

template <typename T>
int launch_nchw_to_nhwc(const float* in, T* out, T* out_lo, int n, int c, int h, int w, cudaStream_t st) {
  const int hw = h * w;
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n), block(32, 8);
  nchw_to_nhwc_kernel<T><<<grid, block, 0, st>>>(in, out, out_lo, c, hw);
  return 1;
}
template int launch_nchw_to_nhwc<float>(const float*, float*, float*, int, int, int, int, cudaStream_t);
template int launch_nchw_to_nhwc<bf16>(const float*, bf16*, bf16*, int, int, int, int, cudaStream_t);

template <typename T>
int launch_nhwc_to_nchw(const T* in, float* out, int n, int c, int h, int w, cudaStream_t st) {
  const int hw = h * w;
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n), block(32, 8);
  nhwc_to_nchw_kernel<T, false, float><<<grid, block, 0, st>>>(in, nullptr, out, c, hw);
  return 1;
}
template int launch_nhwc_to_nchw<float>(const float*, float*, int, int, int, int, cudaStream_t);
template int launch_nhwc_to_nchw<bf16>(const bf16*, float*, int, int, int, int, cudaStream_t);

// bf16 NHWC -> bf16 planes (A-DISTS statistics of the bf16 trunk: exact copies, half the bytes of fp32 planes).
// Tile = 256 pixels x 64 channels per block: all eight 16-byte loads of a thread are issued up front (32 KB per block
// in flight), the tile is transposed through shared memory as 32-bit channel pairs, and each plane receives 512
// contiguous bytes per block (pixel pairs, 4 bytes per lane, 128 bytes per warp store).
// grid = (ceil(hw/256), c/64, n), block = 256.  c % 64 == 0.
constexpr int kPlaneTile = 256;
__global__ void __launch_bounds__(256)
nhwc_to_planes_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int c, int hw) {
  __shared__ uint32_t tile[kPlaneTile][33];   // [pixel][channel pair]
  const int img = blockIdx.z;
  const int p0 = blockIdx.x * kPlaneTile, c0 = blockIdx.y * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bf16* src = in + (size_t)img * hw * c + c0;
  // load: one 16-byte chunk (8 channels) per lane, 4 pixels per warp instruction
  const int lp = lane >> 3, lc = lane & 7;
  uint4 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int p = p0 + (k * 8 + warp) * 4 + lp;
    v[k] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)min(p, hw - 1) * c) + lc);   // clamped, unconditional
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t* row = &tile[(k * 8 + warp) * 4 + lp][lc * 4];
    row[0] = v[k].x; row[1] = v[k].y; row[2] = v[k].z; row[3] = v[k].w;
  }
  __syncthreads();
  bf16* dst = out + (size_t)img * c * hw;
  const bool even = (hw & 1) == 0;     // 4-byte stores need even plane sizes (plane starts stay 4-byte aligned)
  for (int ch = warp; ch < 64; ch += 8) {
    unsigned short* row = reinterpret_cast<unsigned short*>(dst + (size_t)(c0 + ch) * hw);
#pragma unroll
    for (int q = 0; q < kPlaneTile / 64; ++q) {
      const int pp = q * 32 + lane, p = p0 + 2 * pp;
      const uint32_t a = tile[2 * pp][ch >> 1], b = tile[2 * pp + 1][ch >> 1];
      const uint32_t lo = (ch & 1) ? (a >> 16) : (a & 0xffffu), hi = (ch & 1) ? (b >> 16) : (b & 0xffffu);
      if (even && p + 1 < hw) {
        *reinterpret_cast<uint32_t*>(row + p) = lo | (hi << 16);
      } else {
        if (p < hw) row[p] = (unsigned short)lo;
        if (p + 1 < hw) row[p + 1] = (unsigned short)hi;
      }
    }
  }
}

int launch_nhwc_to_nchw_bf16(const bf16* in, bf16* out, int n, int c, int h, int w, cudaStream_t st) {
  const int hw = h * w;
  dim3 grid(ceil_div(hw, kPlaneTile), c / 64, n);
  nhwc_to_planes_bf16_kernel<<<grid, 256, 0, st>>>(in, out, c, hw);
  return 1;
}

int launch_nhwc_split_to_nchw(const bf16* hi, const bf16* lo, float* out, int n, int c, int h, int w, cudaStream_t st) {
  const int hw = h * w;
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n), block(32, 8);
  nhwc_to_nchw_kernel<bf16, true, float><<<grid, block, 0, st>>>(hi, lo, out, c, hw);
  return 1;
}

}  // namespace nqa
