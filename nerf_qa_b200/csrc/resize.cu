// Frame resizing in front of the scorers: the reference resizes every decoded frame on the
// host before scoring -- F.interpolate(bilinear, align_corners=False) in prep.py:91-96 and test2_prep.py:424-439
// ("pixel count" variant), PIL's antialiased bilinear resize in test2_prep.py:109-114 (short side 256) and :221-227
// (256 x 256).  Both filters here, on NCHW fp32 device tensors, so that full-resolution frames can be uploaded once and
// resized where they are scored.
//   antialias = 0: torch upsample_bilinear2d, align_corners=False: src = scale (dst + 0.5) - 0.5 clamped at 0, two taps.
//   antialias = 1: torch _upsample_bilinear2d_aa == PIL ImagingResample with the triangle filter: taps over
//                  [center - support, center + support), support = max(scale, 1), weights normalised per output pixel
//                  (PIL additionally rounds to uint8 after each pass).
#include "common.cuh"
#include "kernels.h"

namespace nqa {
namespace {

__device__ inline void aa_span(int dst, float scale, int in_size, int* xmin, int* xsize, float* center, float* invscale,
                               float* support) {
  // aten/src/ATen/native/cuda/UpSample.cuh _compute_indices_min_size_weights_aa (interp_size = 2, triangle filter)
  *support = scale >= 1.f ? scale : 1.f;
  *invscale = scale >= 1.f ? 1.f / scale : 1.f;
  *center = scale * (dst + 0.5f);
  *xmin = max((int)(*center - *support + 0.5f), 0);
  *xsize = min((int)(*center + *support + 0.5f), in_size) - *xmin;
}
__device__ inline float triangle(float x) {
  x = fabsf(x);
  return x < 1.f ? 1.f - x : 0.f;
}

__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int h, int w, int oh, int ow,
                       float sy, float sx) {
  const size_t total = (size_t)planes * oh * ow;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const int ox = (int)(idx % ow);
    const size_t t = idx / ow;
    const int oy = (int)(t % oh);
    const float* p = in + (t / oh) * (size_t)h * w;
    const float fy = fmaxf(sy * (oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * (ox + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    out[idx] = hy * (hx * __ldg(p + (size_t)y0 * w + x0) + lx * __ldg(p + (size_t)y0 * w + x1)) +
               ly * (hx * __ldg(p + (size_t)y1 * w + x0) + lx * __ldg(p + (size_t)y1 * w + x1));
  }
}

__global__ void __launch_bounds__(256)
resize_bilinear_aa_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int h, int w, int oh, int ow,
                          float sy, float sx) {
  const size_t total = (size_t)planes * oh * ow;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const int ox = (int)(idx % ow);
    const size_t t = idx / ow;
    const int oy = (int)(t % oh);
    const float* p = in + (t / oh) * (size_t)h * w;
    int ymin, ysize, xmin, xsize;
    float cy, iy, supy, cx, ix, supx;
    aa_span(oy, sy, h, &ymin, &ysize, &cy, &iy, &supy);
    aa_span(ox, sx, w, &xmin, &xsize, &cx, &ix, &supx);
    float wxs = 0.f, wys = 0.f;
    for (int j = 0; j < xsize; ++j) wxs += triangle((j + xmin - cx + 0.5f) * ix);
    for (int i = 0; i < ysize; ++i) wys += triangle((i + ymin - cy + 0.5f) * iy);
    // horizontal pass per row, then the vertical combination (the order of the reference's two separable passes)
    float acc = 0.f;
    for (int i = 0; i < ysize; ++i) {
      const float* row = p + (size_t)(ymin + i) * w + xmin;
      float r = 0.f;
      for (int j = 0; j < xsize; ++j) r = fmaf(triangle((j + xmin - cx + 0.5f) * ix) / wxs, __ldg(row + j), r);
      acc = fmaf(triangle((i + ymin - cy + 0.5f) * iy) / wys, r, acc);
    }
    out[idx] = acc;
  }
}

}  // namespace

int launch_resize_bilinear(const float* in, float* out, int planes, int h, int w, int oh, int ow, bool antialias,
                           cudaStream_t st) {
  const size_t total = (size_t)planes * oh * ow;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const float sy = (float)h / (float)oh, sx = (float)w / (float)ow;   // area_pixel_compute_scale, align_corners=False
  if (antialias)
    resize_bilinear_aa_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, planes, h, w, oh, ow, sy, sx);
  else
    resize_bilinear_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, planes, h, w, oh, ow, sy, sx);
  return 1;
}

}  // namespace nqa
