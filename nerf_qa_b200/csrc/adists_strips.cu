// A-DISTS trunk, bf16 path: L2pooling + the level's channel sums from STRIP planes.
//
// When A-DISTS scores are wanted, the last convolution of a VGG stage writes its output as bf16 planes in the strip
// layout [image][channel][w / 8][h][8] (conv_tc_pair.cu / stage1_fused.cu epilogues: a 16 x 8 pixel tile of one channel
// is 256 contiguous bytes) -- the format the tensor-pipe window kernel reads (adists_tc.cu) -- instead of the NHWC tensor
// plus a transposition pass.  This kernel is the level's only other reader: ONE pass over the strip planes of x and y
// gives
//   * the L2-pooled NHWC input of the next stage (DISTS_pt.py:22-25: sqrt(hanning3x3(x^2, stride 2, pad 1) + 1e-12), same
//     operation order as pool_stats_kernel), and
//   * the eight per-channel sums of planes_sums_kernel (adists.cu): x, y, x^2, y^2, xy, relu x, relu(x)^2,
//     relu x log2 relu x (normalisation constants, means and the closed-form entropy of ADISTS.py:127-135,152-167).
//
// Mapping.  Block = 8 warps = one 64-channel slab x `spb` consecutive strips x 31 pooled rows.  A warp owns 8 channels;
// lane l works on pooled row oy = 31 rc + l - 1 of the current strip: it loads the strip rows 2oy and 2oy + 1 of x and y
// (two 16-byte loads per image; consecutive lanes read consecutive 32 bytes), keeps the last column of the previous strip
// for the left tap, and takes the horizontal sums of row 2oy - 1 from lane l - 1 (lane 0 only provides that row: 31 of
// 32 lanes produce output).  The four pooled pixels of a lane go to a padded shared-memory tile [pixel][64 channels] and
// leave as 8-byte pieces of full 128-byte NHWC pixel rows.  Sums: a butterfly per (channel, strip) leaves every total in
// all lanes; lane l accumulates entries l and l + 32 of the warp's 8 x 8 table -- a fixed order that depends on the level
// shape only, so a frame's sums do not depend on the batch it travels in.
// grid = (chunks, c / 64, pairs); partial[((b * chunks + chunk) * 8 + s) * c + ch].
#include "adists.h"
#include "common.cuh"
#include "kernels.h"

namespace nqa {
namespace {

constexpr int kRowsPerChunk = 31;      // pooled rows per block (lane 0 of a warp is the halo lane)
constexpr int kPitch = 136;            // bytes per staged pixel row (64 channels + 8 bytes of padding)
constexpr int kTileBytes = 4 * 32 * kPitch;

__device__ __forceinline__ float2 lo_hi(uint32_t v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
__device__ __forceinline__ float root_approx(float v) {   // rounded to bf16 afterwards (same as pool_stats_kernel)
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ void unpack_row(const uint4& t, float2 (&v)[4]) {
  v[0] = lo_hi(t.x); v[1] = lo_hi(t.y); v[2] = lo_hi(t.z); v[3] = lo_hi(t.w);
}

// horizontal [1 2 1] of squares at the four pooled columns of a strip row: left = the pixel in front of the strip
__device__ __forceinline__ void hsum(const float2 (&v)[4], float left, float (&h)[4]) {
  float prev = left;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float xx1 = v[j].x * v[j].x, xx2 = v[j].y * v[j].y;
    h[j] = fmaf(prev, prev, fmaf(2.f, xx1, xx2));
    prev = v[j].y;
  }
}

__global__ void __launch_bounds__(256)
pool_strips_sums_kernel(const bf16* __restrict__ planes, bf16* __restrict__ pooled_out, int pairs, int c, int h, int w, int ho,
                        int wo, int spb, int n_sg, int chunks, float* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char stage[];   // [image 2][pooled column 4][lane 32][kPitch]
  const int chunk = blockIdx.x, slab = blockIdx.y, b = blockIdx.z;
  const int sg = chunk % n_sg, rc = chunk / n_sg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strips = w >> 3;
  const int s_begin = sg * spb, s_end = min(strips, s_begin + spb);
  const int oy = rc * kRowsPerChunk + lane - 1;
  const bool own = lane > 0 && oy < ho;
  const int r0 = 2 * oy, r1 = 2 * oy + 1;
  const bool ok0 = oy >= 0 && r0 < h, ok1 = oy >= 0 && r1 < h;      // lane 0 of the first chunk stands for row -1: zeros
  const int cr0 = min(max(r0, 0), h - 1), cr1 = min(max(r1, 0), h - 1);
  const size_t plane = (size_t)h * w;
  const size_t strip_stride = (size_t)h * 8;

  float tot0 = 0.f, tot1 = 0.f;   // entries lane and lane + 32 of the warp's [8 channels][8 sums] table
  const bool do_pool = pooled_out != nullptr;   // level 5 has no pooling behind it: sums only

  for (int s = s_begin; s < s_end; ++s) {
#pragma unroll 1
    for (int ci = 0; ci < 8; ++ci) {
      const int ch = slab * 64 + warp * 8 + ci;
      float sum[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) sum[q] = 0.f;
      float2 rows[2][2][4];   // [image][row][pixel pair]
      uint32_t outw[2][2];
#pragma unroll
      for (int im = 0; im < 2; ++im) {
        const bf16* p = planes + ((size_t)(im * pairs + b) * c + ch) * plane + (size_t)s * strip_stride;
        // unconditional loads from clamped rows, zeroed afterwards (loads under branches serialise)
        uint4 t0 = __ldg(reinterpret_cast<const uint4*>(p + (size_t)cr0 * 8));
        uint4 t1 = __ldg(reinterpret_cast<const uint4*>(p + (size_t)cr1 * 8));
        // left tap: last pixel of the previous strip's rows (the image border is zero)
        const bf16* pl = p - strip_stride + 7;
        const unsigned short l0 = s > 0 ? __ldg(reinterpret_cast<const unsigned short*>(pl + (size_t)cr0 * 8)) : (unsigned short)0;
        const unsigned short l1 = s > 0 ? __ldg(reinterpret_cast<const unsigned short*>(pl + (size_t)cr1 * 8)) : (unsigned short)0;
        if (!ok0) t0 = make_uint4(0u, 0u, 0u, 0u);
        if (!ok1) t1 = make_uint4(0u, 0u, 0u, 0u);
        const float left0 = ok0 ? __uint_as_float((uint32_t)l0 << 16) : 0.f;
        const float left1 = ok1 ? __uint_as_float((uint32_t)l1 << 16) : 0.f;
        unpack_row(t0, rows[im][0]);
        unpack_row(t1, rows[im][1]);
        float ha[4], hb[4], hp[4];
        hsum(rows[im][0], left0, ha);
        hsum(rows[im][1], left1, hb);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          hp[j] = __shfl_up_sync(0xffffffffu, hb[j], 1);   // row 2oy - 1 = the odd row of the lane above
          o[j] = root_approx(fmaf(fmaf(2.f, ha[j], hp[j]) + hb[j], 1.f / 16.f, 1e-12f));
        }
        outw[im][0] = Vec<bf16>::pack2(o[0], o[1]);
        outw[im][1] = Vec<bf16>::pack2(o[2], o[3]);
      }
      // the eight sums over the pixels this lane owns (rows 2oy, 2oy + 1 inside the image; zeros add nothing)
      if (own) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 a = rows[0][r][k], d = rows[1][r][k];
            const float2 rl = make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
            const float2 lg = make_float2(__log2f(fmaxf(a.x, 1e-30f)), __log2f(fmaxf(a.y, 1e-30f)));
            sum[0] += a.x + a.y; sum[1] += d.x + d.y;
            sum[2] = fmaf(a.x, a.x, fmaf(a.y, a.y, sum[2]));
            sum[3] = fmaf(d.x, d.x, fmaf(d.y, d.y, sum[3]));
            sum[4] = fmaf(a.x, d.x, fmaf(a.y, d.y, sum[4]));
            sum[5] += rl.x + rl.y;
            sum[6] = fmaf(rl.x, rl.x, fmaf(rl.y, rl.y, sum[6]));
            sum[7] = fmaf(rl.x, lg.x, fmaf(rl.y, lg.y, sum[7]));
          }
        // stage the four pooled pixels of both images: 2-byte stores, pixel index = column * 32 + lane
        const int chl = warp * 8 + ci;
        if (do_pool) {
#pragma unroll
          for (int im = 0; im < 2; ++im)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t wv = outw[im][j >> 1];
              *reinterpret_cast<unsigned short*>(stage + im * kTileBytes + (j * 32 + lane) * kPitch + chl * 2) =
                  (unsigned short)((j & 1) ? (wv >> 16) : (wv & 0xffffu));
            }
        }
      }
      // butterfly: every lane gets the warp total of each sum (lanes that own nothing contribute zeros)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float v = sum[q];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        const float add = ((lane >> 3) == (ci & 3) && (lane & 7) == q) ? v : 0.f;
        if (ci < 4) tot0 += add; else tot1 += add;
      }
    }
    if (!do_pool) continue;
    __syncthreads();
    // write-out: 4 pooled columns x 31 rows x 64 channels per image as 8-byte pieces of 128-byte pixel rows
    for (int i = threadIdx.x; i < 2 * 4 * 32 * 16; i += 256) {
      const int piece = i & 15, j = (i >> 4) & 3, ln = (i >> 6) & 31, im = i >> 11;
      const int oyy = rc * kRowsPerChunk + ln - 1;
      if (ln == 0 || oyy >= ho) continue;
      const uint2 v = *reinterpret_cast<const uint2*>(stage + im * kTileBytes + (j * 32 + ln) * kPitch + piece * 8);
      bf16* dst = pooled_out + (((size_t)(im * pairs + b) * ho + oyy) * wo + (size_t)s * 4 + j) * c + slab * 64 + piece * 4;
      *reinterpret_cast<uint2*>(dst) = v;
    }
    __syncthreads();
  }
  // lane l of warp wq holds entries (ci = l / 8 + 4 t, q = l % 8), t = 0, 1, of the warp's table
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int ci = (lane >> 3) + 4 * t, q = lane & 7;
    partials[(((size_t)b * chunks + chunk) * 8 + q) * c + slab * 64 + warp * 8 + ci] = t ? tot1 : tot0;
  }
}

}  // namespace

bool adists_strips_supported(int level, int H, int W, int window) {
  int h = H, w = W;
  for (int k = 2; k <= level; ++k) { h = pooled(h); w = pooled(w); }
  // the strip layout needs w % 8 == 0 (strips, tensor-map strides); its only readers are this kernel and the tensor-pipe
  // window kernel (windowed levels)
  return level >= 1 && level_channels(level) % 64 == 0 && (w & 7) == 0 && h >= window && w >= window && window >= 3 && window <= 21;
}

void adists_strips_plan(int c, int h, int w, int* spb, int* n_sg, int* chunks) {
  const int strips = w >> 3, rcs = ceil_div(pooled(h), kRowsPerChunk);
  // ~512 chunks per (pair, 64-channel slab) at most: the partition depends on the level shape only
  int per = 1;
  while (ceil_div(strips, per) * rcs > 512) ++per;
  *spb = per;
  *n_sg = ceil_div(strips, per);
  *chunks = *n_sg * rcs;
}

// pooled_out may be null (level 5: no pooling after stage 5, sums only)
int launch_pool_strips_sums(const bf16* planes, bf16* pooled_out, int pairs, int c, int h, int w, float* partials, cudaStream_t st) {
  int spb, n_sg, chunks;
  adists_strips_plan(c, h, w, &spb, &n_sg, &chunks);
  const int smem = 2 * kTileBytes;
  ensure_dynamic_smem(pool_strips_sums_kernel, smem);
  dim3 grid(chunks, c / 64, pairs);
  pool_strips_sums_kernel<<<grid, 256, smem, st>>>(planes, pooled_out, pairs, c, h, w, pooled(h), pooled(w), spb, n_sg, chunks,
                                                  partials);
  return 1;
}

}  // namespace nqa
