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
constexpr int kCarryBytes = 8 * 8 * 32 * 8;   // [warp][channel][lane] last pixels of the previous strip's four rows

__device__ uint4 g_zero_rows;          // what a row outside the image reads as

__device__ __forceinline__ float2 lo_hi(uint32_t v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
__device__ __forceinline__ float root_approx(float v) {   // rounded to bf16 afterwards (same as pool_stats_kernel)
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float lg2_approx(float v) {    // v >= 1e-30: no denormal handling needed
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// horizontal [1 2 1] of squares at the four pooled columns of a strip row from the squares of its pixel pairs;
// left2 = square of the pixel in front of the strip.  Same values as fma(x0, x0, fma(2, x1 x1, x2 x2)): the squares of
// bf16 values are exact in fp32.
__device__ __forceinline__ void hsum(const float2 (&sq)[4], float left2, float (&h)[4]) {
  float prev = left2;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = fmaf(2.f, sq[j].x, sq[j].y) + prev;
    prev = sq[j].y;
  }
}

// Sum of each of eight per-lane values over the warp by recursive halving (9 shuffles instead of 40): afterwards lane l
// holds the total of value l / 4.  Fixed order.
__device__ __forceinline__ float reduce8(float (&v)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float a[4], c[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = b4 ? v[i + 4] : v[i], send = b4 ? v[i] : v[i + 4];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = b3 ? a[i + 2] : a[i], send = b3 ? a[i] : a[i + 2];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float keep = b2 ? c[1] : c[0], send = b2 ? c[0] : c[1];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// kDists: also the five pivot-shifted DISTS moments of the level (x - x0, y - y0, their squares and product; pivot = the
// channel's first pixel; DISTS_pt.py:110-125, layout of pool_stats_kernel: dpartials[((b * chunks + chunk) * 5 + q) * c + ch],
// pivots behind them), for calls that want both metrics from one trunk pass.
template <bool kDists>
__global__ void __launch_bounds__(256, kDists ? 3 : 4)
pool_strips_sums_kernel(const bf16* __restrict__ planes, bf16* __restrict__ pooled_out, int pairs, int c, int h, int w, int ho,
                        int wo, int spb, int n_sg, int chunks, float* __restrict__ partials, float* __restrict__ dpartials) {
  extern __shared__ __align__(16) unsigned char smem[];   // staging [image 2][pooled column 4][lane 32][kPitch] + carry
  unsigned char* stage = smem;
  uint2* carry = reinterpret_cast<uint2*>(smem + 2 * kTileBytes);
  const int chunk = blockIdx.x, slab = blockIdx.y, b = blockIdx.z;
  const int sg = chunk % n_sg, rc = chunk / n_sg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strips = w >> 3;
  const int s_begin = sg * spb, s_end = min(strips, s_begin + spb);
  const int oy = rc * kRowsPerChunk + lane - 1;
  const bool own = lane > 0 && oy < ho;
  const int r0 = 2 * oy, r1 = 2 * oy + 1;
  const bool ok0 = oy >= 0 && r0 < h, ok1 = oy >= 0 && r1 < h;      // lane 0 of the first chunk stands for row -1: zeros
  const size_t plane = (size_t)h * w;
  const size_t strip_stride = (size_t)h * 8;
  const bool do_pool = pooled_out != nullptr;   // level 5 has no pooling behind it: sums only

  float tot0 = 0.f, tot1 = 0.f;   // lane l: totals of sum l / 4 for the channels l % 4 and l % 4 + 4 of the warp
  float dtot0 = 0.f, dtot1 = 0.f;  // the same for the DISTS moments

  // Byte pointers to this lane's two rows of x and y, advanced by one plane per channel (32-bit increment: H W <= 2^28).
  // Rows outside the image point at a zero row and do not advance: no select per loaded word, no load under a branch,
  // and ~10 integer instructions of address arithmetic per channel instead of ~100 for indices rebuilt from scratch.
  const uint32_t plane_b = (uint32_t)(plane * 2);
  const uint32_t inc0 = ok0 ? plane_b : 0u, inc1 = ok1 ? plane_b : 0u;
  const size_t xy_b = (size_t)pairs * c * plane * 2;
  const char* zero = reinterpret_cast<const char*>(&g_zero_rows);
  for (int s = s_begin; s < s_end; ++s) {
    const char* base = reinterpret_cast<const char*>(planes) +
                       (((size_t)b * c + slab * 64 + warp * 8) * plane + (size_t)s * strip_stride) * 2;
    const char* ax0 = ok0 ? base + (size_t)r0 * 16 : zero;
    const char* ax1 = ok1 ? base + (size_t)r1 * 16 : zero;
    const char* ay0 = ok0 ? ax0 + xy_b : zero;
    const char* ay1 = ok1 ? ax1 + xy_b : zero;
#pragma unroll 1
    for (int ci = 0; ci < 8; ++ci, ax0 += inc0, ax1 += inc1, ay0 += inc0, ay1 += inc1) {
      // the next channel's rows on their way into L2 while this one is processed (every other lane: one request per 64
      // bytes; costs no registers, unlike a software pipeline of the loads)
      if ((lane & 1) == 0 && ci < 7) {   // (stays inside the slab: never past the end of the planes)
        prefetch_l2(ax0 + inc0);
        prefetch_l2(ay0 + inc0);
      }
      const uint4 tx0 = __ldg(reinterpret_cast<const uint4*>(ax0));
      const uint4 tx1 = __ldg(reinterpret_cast<const uint4*>(ax1));
      const uint4 ty0 = __ldg(reinterpret_cast<const uint4*>(ay0));
      const uint4 ty1 = __ldg(reinterpret_cast<const uint4*>(ay1));
      // left taps: the last pixels of the previous strip's rows -- from global memory for the first strip of the block,
      // from the carry this lane left in shared memory afterwards (pixels x r0, x r1, y r0, y r1 as bf16)
      uint2 left;
      if (s == s_begin) {
        left = make_uint2(0u, 0u);
        if (s > 0) {   // (a zero row stays a zero row 16 h - 14 bytes earlier only by luck: select)
          const ptrdiff_t back = 14 - (ptrdiff_t)strip_stride * 2;
          const uint32_t a0 = ok0 ? __ldg(reinterpret_cast<const unsigned short*>(ax0 + back)) : 0u;
          const uint32_t a1 = ok1 ? __ldg(reinterpret_cast<const unsigned short*>(ax1 + back)) : 0u;
          const uint32_t d0 = ok0 ? __ldg(reinterpret_cast<const unsigned short*>(ay0 + back)) : 0u;
          const uint32_t d1 = ok1 ? __ldg(reinterpret_cast<const unsigned short*>(ay1 + back)) : 0u;
          left = make_uint2(a0 | (a1 << 16), d0 | (d1 << 16));
        }
      } else {
        left = carry[(warp * 8 + ci) * 32 + lane];
      }
      carry[(warp * 8 + ci) * 32 + lane] = make_uint2((tx0.w >> 16) | (tx1.w & 0xffff0000u), (ty0.w >> 16) | (ty1.w & 0xffff0000u));

      const uint32_t wx[2][4] = {{tx0.x, tx0.y, tx0.z, tx0.w}, {tx1.x, tx1.y, tx1.z, tx1.w}};
      const uint32_t wy[2][4] = {{ty0.x, ty0.y, ty0.z, ty0.w}, {ty1.x, ty1.y, ty1.z, ty1.w}};
      // sums (packed fp32 pairs; the features are ReLU outputs, so relu x = x and its sums are sums 0 and 2):
      // x, y, x^2, y^2, xy, x log2 x
      float2 acc[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[q] = make_float2(0.f, 0.f);
      float2 sqx[2][4], sqy[2][4];
      // DISTS moments: pivots negated; a row outside the image (zeros) is shifted by zero so that it adds nothing
      float2 dacc[5];
      float2 npx[2], npy[2];
      if (kDists) {
        const char* cx = reinterpret_cast<const char*>(planes) + ((size_t)b * c + slab * 64 + warp * 8 + ci) * plane * 2;
        const float x0 = __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(cx)) << 16);
        const float y0 = __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(cx + xy_b)) << 16);
        if (chunk == 0 && s == s_begin && lane == 0) {
          float* piv = dpartials + (size_t)gridDim.z * chunks * 5 * c + (size_t)b * 2 * c + slab * 64 + warp * 8 + ci;
          piv[0] = x0;
          piv[c] = y0;
        }
        npx[0] = ok0 ? make_float2(-x0, -x0) : make_float2(0.f, 0.f);
        npx[1] = ok1 ? make_float2(-x0, -x0) : make_float2(0.f, 0.f);
        npy[0] = ok0 ? make_float2(-y0, -y0) : make_float2(0.f, 0.f);
        npy[1] = ok1 ? make_float2(-y0, -y0) : make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 5; ++q) dacc[q] = make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = lo_hi(wx[r][k]), d = lo_hi(wy[r][k]);
          if (kDists) {
            const float2 dx = __fadd2_rn(a, npx[r]), dy = __fadd2_rn(d, npy[r]);
            dacc[0] = __fadd2_rn(dacc[0], dx);
            dacc[1] = __fadd2_rn(dacc[1], dy);
            dacc[2] = __ffma2_rn(dx, dx, dacc[2]);
            dacc[3] = __ffma2_rn(dy, dy, dacc[3]);
            dacc[4] = __ffma2_rn(dx, dy, dacc[4]);
          }
          sqx[r][k] = __fmul2_rn(a, a);
          sqy[r][k] = __fmul2_rn(d, d);
          const float2 lg = make_float2(lg2_approx(fmaxf(a.x, 1e-30f)), lg2_approx(fmaxf(a.y, 1e-30f)));
          acc[0] = __fadd2_rn(acc[0], a);
          acc[1] = __fadd2_rn(acc[1], d);
          acc[2] = __fadd2_rn(acc[2], sqx[r][k]);
          acc[3] = __fadd2_rn(acc[3], sqy[r][k]);
          acc[4] = __ffma2_rn(a, d, acc[4]);
          acc[5] = __ffma2_rn(a, lg, acc[5]);
        }
      // pooling: horizontal sums of the two rows, the row above from the lane above, sqrt, bf16
      uint32_t outw[2][2];
#pragma unroll
      for (int im = 0; im < 2; ++im) {
        const uint32_t lw = im ? left.y : left.x;
        const float l0 = __uint_as_float(lw << 16), l1 = __uint_as_float(lw & 0xffff0000u);
        float ha[4], hb[4], o[4];
        hsum(im ? sqy[0] : sqx[0], l0 * l0, ha);
        hsum(im ? sqy[1] : sqx[1], l1 * l1, hb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float hp = __shfl_up_sync(0xffffffffu, hb[j], 1);   // row 2oy - 1 = the odd row of the lane above
          o[j] = root_approx(fmaf(fmaf(2.f, ha[j], hp) + hb[j], 1.f / 16.f, 1e-12f));
        }
        outw[im][0] = Vec<bf16>::pack2(o[0], o[1]);
        outw[im][1] = Vec<bf16>::pack2(o[2], o[3]);
      }
      if (own && do_pool) {
        // stage the four pooled pixels of both images: 2-byte stores, pixel index = column * 32 + lane
        const int chl = warp * 8 + ci;
#pragma unroll
        for (int im = 0; im < 2; ++im)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t wv = outw[im][j >> 1];
            *reinterpret_cast<unsigned short*>(stage + im * kTileBytes + (j * 32 + lane) * kPitch + chl * 2) =
                (unsigned short)((j & 1) ? (wv >> 16) : (wv & 0xffffu));
          }
      }
      // per-lane sums (lanes that own nothing contribute zeros) -> warp totals; lane l keeps sum l / 4 of channel l % 4 (+ 4)
      float v[8];
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] = own ? acc[q].x + acc[q].y : 0.f;
      v[6] = v[7] = 0.f;
      const float t = reduce8(v, lane);
      const float add = ((lane & 3) == (ci & 3)) ? t : 0.f;
      if (ci < 4) tot0 += add; else tot1 += add;
      if (kDists) {
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] = own ? dacc[q].x + dacc[q].y : 0.f;
        v[5] = v[6] = v[7] = 0.f;
        const float dt = reduce8(v, lane);
        const float dadd = ((lane & 3) == (ci & 3)) ? dt : 0.f;
        if (ci < 4) dtot0 += dadd; else dtot1 += dadd;
      }
    }
    if (!do_pool) continue;
    __syncthreads();
    // write-out: 4 pooled columns x 31 rows x 64 channels per image as 8-byte pieces of 128-byte pixel rows
    for (int i = threadIdx.x; i < 2 * 4 * 32 * 16; i += 256) {
      const int piece = i & 15, j = (i >> 4) & 3, ln = (i >> 6) & 31, im = i >> 11;
      const int oyy = rc * kRowsPerChunk + ln - 1;
      if (ln == 0 || oyy >= ho) continue;
      const uint2 val = *reinterpret_cast<const uint2*>(stage + im * kTileBytes + (j * 32 + ln) * kPitch + piece * 8);
      bf16* dst = pooled_out + (((size_t)(im * pairs + b) * ho + oyy) * wo + (size_t)s * 4 + j) * c + slab * 64 + piece * 4;
      *reinterpret_cast<uint2*>(dst) = val;
    }
    __syncthreads();
  }
  // lane l: sum q = l / 4 (0..5: x, y, x^2, y^2, xy, x log2 x) of channels l % 4 and l % 4 + 4; the partial layout of
  // planes_sums_kernel has eight sums per channel: 5 = relu x = sum 0, 6 = relu(x)^2 = sum 2, 7 = the entropy sum
  const int q = lane >> 2;
  if (q < 6) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int ch = slab * 64 + warp * 8 + (lane & 3) + 4 * t;
      const float val = t ? tot1 : tot0;
      float* dst = partials + ((size_t)b * chunks + chunk) * 8 * c + ch;
      if (q < 5) dst[(size_t)q * c] = val;
      if (q == 0) dst[(size_t)5 * c] = val;
      if (q == 2) dst[(size_t)6 * c] = val;
      if (q == 5) dst[(size_t)7 * c] = val;
      if (kDists && q < 5) dpartials[(((size_t)b * chunks + chunk) * 5 + q) * c + ch] = t ? dtot1 : dtot0;
    }
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
  // ~16 strip groups per row chunk (the first strip of a group fetches its left taps from global memory: the longer
  // the group, the fewer of those sector-sized 2-byte loads) and at most 512 chunks per (pair, 64-channel slab); the
  // partition depends on the level shape only
  int per = ceil_div(strips, 16);
  while (ceil_div(strips, per) * rcs > 512) ++per;
  *spb = per;
  *n_sg = ceil_div(strips, per);
  *chunks = *n_sg * rcs;
}

// pooled_out may be null (level 5: no pooling after stage 5, sums only)
int launch_pool_strips_sums(const bf16* planes, bf16* pooled_out, int pairs, int c, int h, int w, float* partials,
                            float* dists_partials, cudaStream_t st) {
  int spb, n_sg, chunks;
  adists_strips_plan(c, h, w, &spb, &n_sg, &chunks);
  const int smem = 2 * kTileBytes + kCarryBytes;
  dim3 grid(chunks, c / 64, pairs);
  if (dists_partials) {
    ensure_dynamic_smem(pool_strips_sums_kernel<true>, smem);
    pool_strips_sums_kernel<true><<<grid, 256, smem, st>>>(planes, pooled_out, pairs, c, h, w, pooled(h), pooled(w), spb, n_sg, chunks,
                                                          partials, dists_partials);
  } else {
    ensure_dynamic_smem(pool_strips_sums_kernel<false>, smem);
    pool_strips_sums_kernel<false><<<grid, 256, smem, st>>>(planes, pooled_out, pairs, c, h, w, pooled(h), pooled(w), spb, n_sg, chunks,
                                                           partials, nullptr);
  }
  return 1;
}

int adists_strips_chunks(int c, int h, int w) {
  int spb, n_sg, chunks;
  adists_strips_plan(c, h, w, &spb, &n_sg, &chunks);
  return chunks;
}

}  // namespace nqa
