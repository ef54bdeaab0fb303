// L2pooling + per-level DISTS moments for bf16 NHWC levels, fed by TMA.
//
// Same arithmetic, thread mapping, partition and partial-sum layout as pool_stats_kernel<bf16> (pool_stats.cu) -- the
// results are bit-identical -- but the input rows do not travel through registers: one thread issues
// cp.async.bulk.tensor box loads of the next row pairs ({64 channels, 2*PL+1 pixels, 2 rows} per image and 64-channel
// chunk) into a 4-stage shared-memory ring while the block works on the current pair.  The register-fed kernel keeps
// at most 6 x 8 bytes per thread in flight (~37 KB per SM at 3 blocks, ncu: 45% issue-active, 67-71% of the HBM copy
// peak); the ring keeps three full stages per block (~100 KB per SM at 2 blocks) in flight without costing registers.  Image
// borders (column -1, rows -1 / h) are zero-filled by the TMA unit.
//
// Reference semantics: DISTS_pt.py:11-25 (L2pooling), :110-125 (moments).
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace nqa {
namespace pst {

using namespace tc;

constexpr int STAGES = 4;

__device__ __forceinline__ void unpack4(uint2 t, float2 (&v)[2]) {
  v[0] = make_float2(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u));
  v[1] = make_float2(__uint_as_float(t.y << 16), __uint_as_float(t.y & 0xffff0000u));
}
__device__ __forceinline__ float root(float v) {   // rounded to bf16 afterwards: the 1-ulp MUFU square root is enough
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

__global__ void __launch_bounds__(256, 2)
pool_stats_tma_kernel(const __grid_constant__ CUtensorMap tmap, const bf16* __restrict__ fx, const bf16* __restrict__ fy,
                      bf16* __restrict__ px_out, bf16* __restrict__ py_out, int pairs, int h, int w, int c, int ho, int wo,
                      int R, int chunks, int box_bytes, float* __restrict__ partials) {
  constexpr int V = 4;
  const int G = c / V, PL = 256 / G, BW = 2 * PL + 1, nchunk = c / 64;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int b = blockIdx.z;
  const int ox = blockIdx.x * PL + pl;
  const int oy0 = blockIdx.y * R, oy1 = min(ho, oy0 + R);
  const bool active = ox < wo;
  const float2 zero2 = make_float2(0.f, 0.f), two2 = make_float2(2.f, 2.f);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int stage_bytes = (2 * nchunk * box_bytes + 1023) & ~1023;   // 1024-aligned stages: same swizzle phases in each
  __shared__ uint64_t full[STAGES];
  __shared__ float red[256 * V];

  // iteration j consumes input rows 2(oy0 + j - 1) and 2(oy0 + j - 1) + 1; j = 0 only provides the row above the strip
  const int n_it = oy1 - oy0 + 1;
  const int x_start = 2 * (int)blockIdx.x * PL - 1;
  auto issue = [&](int j) {   // thread 0
    const int s = j % STAGES;
    uint8_t* dst = ring + s * stage_bytes;
    mbar_expect_tx(&full[s], (uint32_t)(2 * nchunk * 128 * BW * 2));
    for (int t = 0; t < 2; ++t)
      for (int q = 0; q < nchunk; ++q)
        tma_load_4d(dst + (t * nchunk + q) * box_bytes, &tmap, &full[s], q * 64, x_start, 2 * (oy0 + j - 1), b + t * pairs);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int j = 0; j < STAGES && j < n_it; ++j) issue(j);
  }
  __syncthreads();

  float2 s[5][2];
#pragma unroll
  for (int q = 0; q < 5; ++q) s[q][0] = s[q][1] = zero2;

  const size_t ioff = (size_t)b * h * w * c + (size_t)g * V;
  const size_t ooff = (size_t)b * ho * wo * c + (size_t)g * V;
  const bool c2 = 2 * ox + 1 < w;   // column 2ox always exists for an active thread
  // pivots = the channel's first pixel, negated (kernels.h: shifted one-pass sums)
  float2 npx[2] = {zero2, zero2}, npy[2] = {zero2, zero2};
  if (active) {
    unpack4(__ldg(reinterpret_cast<const uint2*>(fx + ioff)), npx);
    unpack4(__ldg(reinterpret_cast<const uint2*>(fy + ioff)), npy);
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
  }

  // this thread's 8 bytes inside a 128-byte pixel row of its 64-channel chunk (128-byte swizzle: chunk ^ (row & 7))
  const int q64 = g >> 4, j16 = (g & 15) >> 1, half8 = (g & 1) * 8;
  const uint32_t box_x = (uint32_t)(q64 * box_bytes), box_y = (uint32_t)((nchunk + q64) * box_bytes);
  // Boxes are packed at 128-byte granularity inside 1024-aligned stages: the swizzle phase of a pixel row is bits 7..9
  // of its offset in the stage (what the TMA unit XORs into bits 4..6 of the ABSOLUTE address).  The twelve offsets a
  // thread reads per row pair (2 images x 2 rows x 3 columns) do not depend on the stage: computed once.
  uint32_t off_x[2][3], off_y[2][3];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint32_t ax_ = box_x + (uint32_t)(r * BW + 2 * pl + k) * 128, ay_ = box_y + (uint32_t)(r * BW + 2 * pl + k) * 128;
      off_x[r][k] = ax_ + ((j16 ^ ((ax_ >> 7) & 7)) << 4) + half8;
      off_y[r][k] = ay_ + ((j16 ^ ((ay_ >> 7) & 7)) << 4) + half8;
    }
  auto lds8 = [&](const uint8_t* stage, uint32_t off) { return *reinterpret_cast<const uint2*>(stage + off); };

  // horizontal [1 2 1] of squares of one staged input row for both images (+ statistics on the owned columns)
  auto row = [&](const uint8_t* stage, int r, bool in_image, bool own, float2 (&hx)[2], float2 (&hy)[2]) {
    if (!in_image) {
      hx[0] = hx[1] = hy[0] = hy[1] = zero2;
      return;
    }
    float2 x0[2], x1[2], x2[2], y0[2], y1[2], y2[2];
    unpack4(lds8(stage, off_x[r][0]), x0);
    unpack4(lds8(stage, off_x[r][1]), x1);
    unpack4(lds8(stage, off_x[r][2]), x2);
    unpack4(lds8(stage, off_y[r][0]), y0);
    unpack4(lds8(stage, off_y[r][1]), y1);
    unpack4(lds8(stage, off_y[r][2]), y2);
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

  bf16* ox_ptr = px_out + ooff + ((size_t)oy0 * wo + ox) * c;
  bf16* oy_ptr = py_out + ooff + ((size_t)oy0 * wo + ox) * c;
  const ptrdiff_t ostride = (ptrdiff_t)wo * c;
  float2 px[2] = {zero2, zero2}, py[2] = {zero2, zero2};   // horizontal sums of the odd row above (2oy - 1)
  const float2 sixteenth = make_float2(1.f / 16.f, 1.f / 16.f), eps = make_float2(1e-12f, 1e-12f);
#pragma unroll 1
  for (int j = 0; j < n_it; ++j) {
    const int st = j % STAGES;                              // STAGES is a power of two
    const uint8_t* stage = ring + st * stage_bytes;
    mbar_wait(&full[st], (uint32_t)((j / STAGES) & 1));
    if (active) {
      // j = 0: only the horizontal sums of the row above the strip (no statistics, no output)
      const bool first = j == 0;
      const int oy = oy0 + j - 1;
      float2 ax[2], ay[2], bx2[2], by2[2];
      row(stage, 0, !first, true, ax, ay);                   // row 2oy always exists
      row(stage, 1, first ? oy0 > 0 : 2 * oy + 1 < h, !first, bx2, by2);
      if (!first) {
        uint2 tx, ty;
        float o[8];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float2 vx = __ffma2_rn(__fadd2_rn(__ffma2_rn(two2, ax[i], px[i]), bx2[i]), sixteenth, eps);
          const float2 vy = __ffma2_rn(__fadd2_rn(__ffma2_rn(two2, ay[i], py[i]), by2[i]), sixteenth, eps);
          o[2 * i] = root(vx.x); o[2 * i + 1] = root(vx.y);
          o[4 + 2 * i] = root(vy.x); o[4 + 2 * i + 1] = root(vy.y);
        }
        tx.x = Vec<bf16>::pack2(o[0], o[1]); tx.y = Vec<bf16>::pack2(o[2], o[3]);
        ty.x = Vec<bf16>::pack2(o[4], o[5]); ty.y = Vec<bf16>::pack2(o[6], o[7]);
        *reinterpret_cast<uint2*>(ox_ptr) = tx;
        *reinterpret_cast<uint2*>(oy_ptr) = ty;
        ox_ptr += ostride;
        oy_ptr += ostride;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        px[i] = bx2[i];
        py[i] = by2[i];
      }
    }
    __syncthreads();                                        // every thread has read this stage
    if (threadIdx.x == 0 && j + STAGES < n_it) issue(j + STAGES);
  }

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

}  // namespace pst

// Returns the number of launches, 0 if this configuration is not served (the caller falls back to pool_stats_kernel).
int launch_pool_stats_tma(const bf16* fx, const bf16* fy, bf16* px_out, bf16* py_out, int pairs, int h, int w, int c, int chunks,
                          float* partials, cudaStream_t st) {
  using namespace pst;
  if (c % 64 != 0 || c > 1024 || fy != fx + (size_t)pairs * h * w * c) return 0;
  if ((reinterpret_cast<uintptr_t>(fx) & 15) != 0) return 0;
  int r, cb, strips;
  pool_stats_plan(h, w, c, 2, &r, &cb, &strips);
  if (cb * strips != chunks) return 0;
  const int pl = 256 / (c / 4), bw = 2 * pl + 1;
  const int box_bytes = 128 * bw * 2;
  CUtensorMap map;
  if (conv_tc_make_act_map(&map, fx, 2 * pairs, h, w, c, bw, 2) != 0) return 0;
  const int smem = STAGES * ((2 * (c / 64) * box_bytes + 1023) & ~1023) + 1024;
  ensure_dynamic_smem(pool_stats_tma_kernel, smem);
  dim3 grid(cb, strips, pairs);
  pool_stats_tma_kernel<<<grid, 256, smem, st>>>(map, fx, fy, px_out, py_out, pairs, h, w, c, pooled(h), pooled(w), r, chunks,
                                                 box_bytes, partials);
  return 1;
}

}  // namespace nqa
