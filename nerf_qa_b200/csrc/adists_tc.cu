// A-DISTS local-window statistics on the tensor pipe (bf16 feature planes of the bf16 trunk): the two 21-tap gaussian
// passes of ADISTS.forward / compute_prob (nerf_qa/ADISTS/ADISTS.py:75-80, 165-175) as banded-Toeplitz matrix products.
//
// Why warp-level mma.sync (HMMA m16n8k16) and not tcgen05 here: the filter matrix is a 21-wide band.  With an M = 16
// output block the band fills 21 of 48 K columns (44%) in the horizontal pass and 21 of 32 (66%) in the vertical one; a
// tcgen05 tile (M = 128 outputs, K = 148 -> 160 inputs) fills 13%, so its 4x higher peak (measured on this pool:
// mma.sync 279 TFMA/s, cuBLAS tcgen05 ~800 TFMA/s) buys nothing, while the accumulator would have to leave TMEM for the
// bf16 hi/lo conversion between the passes.  With mma.sync the first pass' accumulator fragment IS the second pass' A
// fragment (same register layout), so the two passes chain in registers.
//
// Precision.  The features are exact bf16 values, so x^2, y^2 and x*y have <= 16 significant bits and split EXACTLY into
// bf16 hi + lo; the fp32 results of the first pass are split into hi + lo as well (2^-17 relative); the gaussian taps are
// rounded to bf16 once (the same perturbed window for all five maps) and the exact sum of the rounded taps renormalises
// the results.  Variances therefore cancel at ~1e-5 of E[x^2] (fp32 CUDA-core kernel: 1e-7), far inside the 2e-3
// tolerance of the bf16 path; the fp32 / bf16x3 trunks keep the CUDA-core kernel (adists.cu).
//
// Block = 12 warps = 4 (columns) x 3 (rows); warp tile = 16 output columns x 24 output rows; block tile 64 x 72 outputs
// from a 96 x 96 input tile.  Per channel: the raw x / y tiles arrive by TMA (three 32-column SWIZZLE_64B boxes per tile,
// zero-filled outside the image, double-buffered on two mbarriers; one thread issues them), the producer warps write the
// product tiles (x^2, y^2, xy, rounded or as hi / lo; packed bf16 multiplies) once, then every warp runs, per map q in
// (x, y, x^2, y^2, xy):
//   pass 1:  D1[16 out cols][8 rows] += T[16 x 16] * tile[16 cols x 8 rows]      3 k-steps x 6 row blocks x terms
//   split :  D1 -> bf16 hi / lo A fragments (in registers)
//   pass 2:  D2[16 out cols][8 out rows] += D1[16 x 16 rows] * T2[16 rows x 8]   2 k-steps x 3 row blocks x 2 terms
// and folds the five maps into T, S and gamma per pixel, accumulated over the channel loop in registers.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "adists.h"
#include "adists_tc.h"
#include "common.cuh"
#include "kernels.h"

namespace nqa {
namespace {

constexpr int kWarpsX = 4, kOB = 3;                               // warp columns, output row blocks (of 8) per warp
constexpr int kTileOW = 16 * kWarpsX;                             // 64 output columns per block
constexpr int kInW = kTileOW + 32;                                // 96 input columns (halo of 20 rounded up to blocks)
constexpr int kParts = kInW / 32;                                 // a tile = three 32-column parts of kInH rows x 64 bytes,
                                                                  // each in the SWIZZLE_64B layout TMA writes: the 16-byte
                                                                  // chunk q of row r sits at chunk q ^ ((r >> 1) & 3), which
                                                                  // makes the 8-row ldmatrix reads conflict-free
constexpr int kRB = kOB + 3;                                      // input row blocks (of 8) per warp
constexpr int kMaxCPer = 512;                                     // channels per block (per-channel scalars in smem)
// kWY warp rows: 3 = one 12-warp block per SM, 64 x 72 outputs, ten (seven for the one-term kernel) 20 KB tiles.
// Measured alternative for the one-term kernel, kWY = 2 with two 8-warp blocks per SM (seven 15 KB tiles each): the 128
// registers per thread it leaves spill 0.5 KB per thread and the level-1 launch is 13% slower (68.6 vs 60.5 ms of window
// time per 32-pair step), so both kernels run with kWY = 3.
template <int kWY, bool kTwo>
struct TcCfg {
  static constexpr int kThreads = 32 * kWarpsX * kWY;
  static constexpr int kTileOH = 8 * kOB * kWY;
  static constexpr int kInH = kTileOH + 24;
  static constexpr int kPartBytes = kInH * 64;
  static constexpr int kTileBytes = kParts * kPartBytes;
  static constexpr int kProdTiles = kTwo ? 6 : 3;
  static constexpr int kTilesBytes = (4 + kProdTiles) * kTileBytes;  // raw x/y double-buffered + the product tiles
  static constexpr int kSmemBytes = kTilesBytes + 3 * kMaxCPer * 4 + 64 + 1024;   // + scalars, mbarriers, alignment
  static constexpr int kBlocksPerSM = kWY == 2 ? 2 : 1;
};

struct TcTaps {
  float g[32];      // bf16-representable taps, zero beyond the window
  float inv_sum2;   // 1 / (sum of the rounded taps)^2
};

__device__ inline uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ inline float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ inline float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ inline uint32_t pack_bf16x2(float2 v) { return pack_bf16x2(v.x, v.y); }
// p = hi + lo with hi = bf16(p), lo = bf16(p - hi), two values at a time: the residual is one packed FMA, p - 1 * hi
// (exact whenever p - hi is representable, which it is for the products of two bf16 values)
__device__ inline void split_bf16x2(float2 p, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(p);
  lo = pack_bf16x2(__ffma2_rn(make_float2(bf16_lo(hi), bf16_hi(hi)), make_float2(-1.f, -1.f), p));
}

__device__ inline void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// m16n8k8: the last k-block of the first pass only reaches 4 columns into its 16 (window 21: 36 input columns per 16
// outputs), so its upper half is all zero taps
__device__ inline void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
__device__ inline void ldmatrix_x1(uint32_t& r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r) : "r"(addr));
}
__device__ inline void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ inline void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// lean mbarrier wait (no printf: this kernel has no registers to spare); traps instead of hanging the GPU
__device__ inline void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spins > (1u << 26)) __trap();
  }
}
__device__ inline bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ inline float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// num / den for denominators in the normal range (all of them carry a 1e-6 or larger stabiliser): MUFU.RCP + FMUL
// instead of the five instructions of div.approx with its range check
__device__ inline float2 div2_approx(float2 num, float2 den) {
  return __fmul2_rn(num, make_float2(rcp_approx(den.x), rcp_approx(den.y)));
}

// One filtered map of the warp tile: pass 1 over the NT term tiles, hi/lo split, pass 2.  out[ob][4] (raw sums, to be
// scaled by 1/sum(g')^2): thread (lane) holds columns lane/4 (+8), rows 8*ob + 2*(lane%4) + {0,1}.
// kTwo = false: the first-pass results go into the second pass as ONE bf16 term (levels whose maps are large enough
// that the 1e-6 stabilisers of T and S dwarf the rounding, see launch_window_mma).
template <int NT, bool kTwo, int kRbStride>
__device__ inline void filtered_map(const uint32_t (&tile_addr)[NT], uint32_t lane_off, uint32_t lane_off2,
                                    const uint32_t (&ta)[3][4], const uint32_t (&tb)[2][2], float (&out)[kOB][4]) {
  uint32_t fh[kRB][2], fl[kTwo ? kRB : 1][2];
#pragma unroll
  for (int rb = 0; rb < kRB; ++rb) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      uint32_t b01[4], b2;
      ldmatrix_x4(b01, tile_addr[t] + lane_off + rb * kRbStride);    // k-blocks 0 and 1 (columns +0, +8, +16, +24)
      ldmatrix_x1(b2, tile_addr[t] + lane_off2 + rb * kRbStride);    // first half of k-block 2 (columns +32 .. +39)
      mma16816(c, ta[0][0], ta[0][1], ta[0][2], ta[0][3], b01[0], b01[1]);
      mma16816(c, ta[1][0], ta[1][1], ta[1][2], ta[1][3], b01[2], b01[3]);
      mma1688(c, ta[2][0], ta[2][1], b2);
    }
    // fp32 -> bf16 hi + lo (round to nearest both; hi + lo carries 16 significant bits)
    if (kTwo) {
      split_bf16x2(make_float2(c[0], c[1]), fh[rb][0], fl[rb][0]);
      split_bf16x2(make_float2(c[2], c[3]), fh[rb][1], fl[rb][1]);
    } else {
      fh[rb][0] = pack_bf16x2(c[0], c[1]);
      fh[rb][1] = pack_bf16x2(c[2], c[3]);
    }
  }
#pragma unroll
  for (int ob = 0; ob < kOB; ++ob) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int rb = ob + 2 * s;
      mma16816(d, fh[rb][0], fh[rb][1], fh[rb + 1][0], fh[rb + 1][1], tb[s][0], tb[s][1]);
      if (kTwo) mma16816(d, fl[rb][0], fl[rb][1], fl[rb + 1][0], fl[rb + 1][1], tb[s][0], tb[s][1]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[ob][i] = d[i];
  }
}

// kStrips: the planes are in the strip layout the trunk's convolution epilogues write for A-DISTS ([plane][w / 8][h][8]:
// 8-pixel-wide column strips, rows contiguous inside a strip; see conv_tc_pair.cu / stage1_fused.cu).  A tile then is
// [row block of 8][strip][8 rows][16 bytes], loaded as one {64 elements, 12 strips} box per row block (each warp issues
// its own row block), and an 8-row ldmatrix read is 128 contiguous bytes.
template <bool kTwo, int kWY, bool kStrips>
__global__ void __launch_bounds__(TcCfg<kWY, kTwo>::kThreads, TcCfg<kWY, kTwo>::kBlocksPerSM)
window_mma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y, int c, int oh, int ow,
                  int tiles_x, int csplit, int choff, const double* __restrict__ chan, const float* __restrict__ wt, TcTaps taps,
                  float* __restrict__ amap, float* __restrict__ bmap, float* __restrict__ gmap) {
  using Cfg = TcCfg<kWY, kTwo>;
  constexpr int kThreads = Cfg::kThreads, kTileOH = Cfg::kTileOH, kTileBytes = Cfg::kTileBytes;
  constexpr int kPartBytes = Cfg::kPartBytes;
  constexpr int kStep = kTwo ? 1 : 2;   // product tile m of (xx_h, xx_l, yy_h, yy_l, xy_h, xy_l) lives in slot 4 + m / kStep
  constexpr int kRowBlocks = Cfg::kInH / 8, kStripRow = (kInW / 8) * 128;   // strip layout: bytes per 8-row block
  constexpr int kRbStride = kStrips ? kStripRow : 8 * 64;
  static_assert(!kStrips || kRowBlocks == kThreads / 32, "strip layout: one row block of the tile per warp");
  extern __shared__ unsigned char smem_dyn[];
  // the swizzle pattern is a function of the shared-memory ADDRESS: tiles start on a 1024-byte boundary (aligned by
  // offset from the __shared__ array so that accesses stay LDS / STS)
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_dyn);
  unsigned char* smem_raw = smem_dyn + ((1024 - (smem_base & 1023)) & 1023);
  const uint32_t smem = (uint32_t)__cvta_generic_to_shared(smem_raw);
  float* s_ax = reinterpret_cast<float*>(smem_raw + Cfg::kTilesBytes);   // per-channel scalars of this block's channels
  float* s_ay = s_ax + kMaxCPer;
  float* s_wc = s_ay + kMaxCPer;
  const uint32_t full_bar = smem + Cfg::kTilesBytes + 3 * kMaxCPer * 4;   // two mbarriers: raw[buf] has landed
  // tiles: raw[buf][x|y] at (buf*2 + i), products xx_h, xx_l, yy_h, yy_l, xy_h, xy_l at 4..9 (kTwo) or xx, yy, xy
  // rounded to bf16 at 4, 5, 6
  auto tile = [&](int i) { return smem + (uint32_t)i * kTileBytes; };

  const int tile_id = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
  const int oy0 = (tile_id / tiles_x) * kTileOH, ox0 = (tile_id % tiles_x) * kTileOW;
  const int c_per = ceil_div(c, csplit);
  const int c_begin = split * c_per, c_end = min(c, c_begin + c_per);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // the same value, known to be warp-uniform
  const int wx = warp % kWarpsX, wy = warp / kWarpsX;

  // Toeplitz fragments (constant over the kernel).  Pass 1, A operand (16 output columns x 16 input columns of k-block
  // d): element (j, k) = g'[16 d + k - j]; pass 2, B operand (16 input rows of k-step s x 8 output rows): g'[16 s + k - i].
  auto tap = [&](int i) { return (i >= 0 && i < 32) ? taps.g[i] : 0.f; };
  uint32_t ta[3][4], tb[2][2];
  {
    const int j0 = lane >> 2, k0 = 2 * (lane & 3);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      ta[d][0] = pack_bf16x2(tap(16 * d + k0 - j0), tap(16 * d + k0 + 1 - j0));
      ta[d][1] = pack_bf16x2(tap(16 * d + k0 - j0 - 8), tap(16 * d + k0 + 1 - j0 - 8));
      ta[d][2] = pack_bf16x2(tap(16 * d + k0 + 8 - j0), tap(16 * d + k0 + 9 - j0));
      ta[d][3] = pack_bf16x2(tap(16 * d + k0 - j0), tap(16 * d + k0 + 1 - j0));   // (k + 8) - (j + 8)
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      tb[s][0] = pack_bf16x2(tap(16 * s + k0 - j0), tap(16 * s + k0 + 1 - j0));
      tb[s][1] = pack_bf16x2(tap(16 * s + k0 + 8 - j0), tap(16 * s + k0 + 9 - j0));
    }
  }

  // raw tile loader (ONE thread): the x and y tiles of channel ch as 2 x 3 TMA boxes of 32 columns x kInH rows; parts of
  // a box outside the image arrive as zeros
  // (strip layout: called by every thread after a block barrier; thread 0 announces the bytes, lane 0 of warp j loads
  // row block j of both tiles)
  auto load_raw = [&](int ch, int buf) {
    const uint32_t bar = full_bar + 8 * buf;
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * kTileBytes) : "memory");
    if (kStrips) {
      // executed by the whole warp with warp-uniform operands, only the instruction predicated on the elected lane
      // (inside an `if (lane == 0)` region every UTMALDG is wrapped in an R2UR + ELECT + BRA.U.ANY loop)
      const uint32_t dst = tile(buf * 2) + warp_u * kStripRow;
      const int row0 = (oy0 + 8 * warp_u) * 8, strip0 = ox0 >> 3, pl = b * c + ch;
      if (elect_one_sync()) {
        tma_load_3d(dst, &map_x, bar, row0, strip0, pl);
        tma_load_3d(dst + kTileBytes, &map_y, bar, row0, strip0, pl);
      }
    } else if (warp_u == 0) {
      const uint32_t dst = tile(buf * 2);
      const int pl = b * c + ch;
      if (elect_one_sync()) {
#pragma unroll
        for (int part = 0; part < kParts; ++part) {
          tma_load_3d(dst + part * kPartBytes, &map_x, bar, ox0 + 32 * part, oy0, pl);
          tma_load_3d(dst + kTileBytes + part * kPartBytes, &map_y, bar, ox0 + 32 * part, oy0, pl);
        }
      }
    }
  };

  // per-lane part of the ldmatrix addresses inside the warp's 48 x 48 input window: row (lane % 8) of chunk
  // 2 wx + lane / 8 (x4: columns +0 .. +24) and of chunk 2 wx + 4 + (lane / 8) % 2 (x2: columns +32, +40); row blocks
  // start at multiples of 8 rows, so the swizzle term (row >> 1) & 3 is (lane >> 1) & 3 for every block
  const uint32_t row_off = (uint32_t)((wy * 8 * kOB + (lane & 7)) * 64), sw = (uint32_t)((lane >> 1) & 3);
  const uint32_t q4 = (uint32_t)(wx * 2 + (lane >> 3)), q2 = (uint32_t)(wx * 2 + 4 + ((lane >> 3) & 1));
  const uint32_t strip_row = (uint32_t)(wy * kOB * kStripRow + (lane & 7) * 16);
  const uint32_t lane_off = kStrips ? strip_row + q4 * 128 : (q4 >> 2) * kPartBytes + row_off + (((q4 & 3) ^ sw) << 4);
  const uint32_t lane_off2 = kStrips ? strip_row + q2 * 128 : (q2 >> 2) * kPartBytes + row_off + (((q2 & 3) ^ sw) << 4);
  const bool warp_live = (oy0 + wy * 8 * kOB) < oh && (ox0 + wx * 16) < ow;
  const bool producer = wy != 0;
  // named barrier 1: every thread arrives once per channel, the producers after writing the product tiles, the row-0
  // warps before they read them
  auto products_ready = [] { asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kThreads) : "memory"); };

  float2 acc_a[kOB][2], acc_b[kOB][2], acc_g[kOB][2];   // element i: pixels (row pair) of columns lane/4 + 8 i
#pragma unroll
  for (int ob = 0; ob < kOB; ++ob)
#pragma unroll
    for (int i = 0; i < 2; ++i) acc_a[ob][i] = acc_b[ob][i] = acc_g[ob][i] = make_float2(0.f, 0.f);

  // per-channel scalars of this block's channels -> shared memory once (1/||x||, 1/||y||, entropy weight)
  for (int i = threadIdx.x; i < c_end - c_begin; i += kThreads) {
    const double* cs = chan + ((size_t)b * kTotalChannels + choff + c_begin + i) * 10;
    s_ax[i] = (float)__ldg(cs);
    s_ay[i] = (float)__ldg(cs + 1);
    s_wc[i] = __ldg(wt + (size_t)b * kTotalChannels + choff + c_begin + i);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_bar) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_bar + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();   // barriers initialised before any warp's TMA signals them
  if (c_begin < c_end) load_raw(c_begin, 0);
  int buf = 0;
  uint32_t phase = 0;   // parity of raw[buf]'s barrier: flips every second channel
  for (int ch = c_begin; ch < c_end; ++ch, buf ^= 1, phase ^= (uint32_t)(buf == 0)) {
    __syncthreads();   // every warp is done with the previous channel's tiles (first pass: barriers + scalars visible)
    if (ch + 1 < c_end) load_raw(ch + 1, buf ^ 1);
    mbar_wait_parity(full_bar + 8 * buf, phase);   // raw[buf] has landed

    // ---- products: x^2, y^2, xy as exact bf16 hi + lo, written once per block by the warps of rows 1 and 2, while the
    // warps of row 0 already filter the two maps that only need the raw tiles.  The three warps of a scheduler (same
    // wx) thereby run out of phase: one issues its MMA bursts while the others do the scalar work between theirs. ----
    if (producer) {
      // The products are elementwise, so the producers walk the tile's 16-byte chunks in PHYSICAL order (chunk tid + 256 k)
      // whatever the swizzle; all loads of the thread are issued before the arithmetic (the shared-memory latency is paid
      // once).  Packed bf16 arithmetic: hi = x * y rounded to bf16 (one HMUL2 per element pair); lo = fma(x, y, -hi), the
      // exact residual (it has <= 8 significant bits).
      constexpr int kProducers = kThreads - 32 * kWarpsX;                           // 256
      constexpr int kIters = (kTileBytes / 16 + kProducers - 1) / kProducers;      // 5
      const int tid = threadIdx.x - 32 * kWarpsX;
      uint4 vx[kIters], vy[kIters];
      uint32_t off[kIters];
#pragma unroll
      for (int k = 0; k < kIters; ++k) {
        off[k] = (uint32_t)(tid + kProducers * k) * 16;
        if (off[k] < (uint32_t)kTileBytes) {
          vx[k] = *reinterpret_cast<const uint4*>(smem_raw + (buf * 2 + 0) * kTileBytes + off[k]);
          vy[k] = *reinterpret_cast<const uint4*>(smem_raw + (buf * 2 + 1) * kTileBytes + off[k]);
        } else {
          off[k] = 0xffffffffu;
        }
      }
#pragma unroll
      for (int k = 0; k < kIters; ++k) {
        if (off[k] == 0xffffffffu) continue;
        const uint32_t xs[4] = {vx[k].x, vx[k].y, vx[k].z, vx[k].w}, ys[4] = {vy[k].x, vy[k].y, vy[k].z, vy[k].w};
        uint32_t o[6][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&xs[e]);
          const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(&ys[e]);
          const __nv_bfloat162 xx = __hmul2(x, x), yy = __hmul2(y, y), xy = __hmul2(x, y);
          o[0][e] = *reinterpret_cast<const uint32_t*>(&xx);
          o[2][e] = *reinterpret_cast<const uint32_t*>(&yy);
          o[4][e] = *reinterpret_cast<const uint32_t*>(&xy);
          if (kTwo) {
            const __nv_bfloat162 xl = __hfma2(x, x, __hneg2(xx)), yl = __hfma2(y, y, __hneg2(yy)), zl = __hfma2(x, y, __hneg2(xy));
            o[1][e] = *reinterpret_cast<const uint32_t*>(&xl);
            o[3][e] = *reinterpret_cast<const uint32_t*>(&yl);
            o[5][e] = *reinterpret_cast<const uint32_t*>(&zl);
          }
        }
#pragma unroll
        for (int m = 0; m < 6; m += (kTwo ? 1 : 2))
          *reinterpret_cast<uint4*>(smem_raw + (4 + m / kStep) * kTileBytes + off[k]) = make_uint4(o[m][0], o[m][1], o[m][2], o[m][3]);
      }
    }
    if (producer) products_ready();
    if (!warp_live) {
      if (!producer) products_ready();
      continue;
    }
    const float ax = s_ax[ch - c_begin], ay = s_ay[ch - c_begin], wc = s_wc[ch - c_begin];
    // maps arrive as raw sums D; the 1 / sum(g')^2 renormalisation and the per-channel 1/||f|| scales are folded
    const float inv = taps.inv_sum2;
    const float2 kx = make_float2(ax * inv, ax * inv), ky = make_float2(ay * inv, ay * inv);
    const float2 kxx = make_float2(ax * ax * inv, ax * ax * inv), kyy = make_float2(ay * ay * inv, ay * ay * inv);
    const float2 kxy = make_float2(ax * ay * inv, ax * ay * inv);
    const float2 c6 = make_float2(1e-6f, 1e-6f), two = make_float2(2.f, 2.f), wc2 = make_float2(wc, wc);
    // gamma = (E[x^2] - m^2) / (m + 1e-12) on the RAW features (ADISTS.py:81-84) = vx / (ax mx + ax^2 1e-12) in
    // normalised units
    const float2 gk = make_float2(ax, ax), ge = make_float2(ax * ax * 1e-12f, ax * ax * 1e-12f);

    float2 mx[kOB][2], my[kOB][2], vs[kOB][2];
    float d[kOB][4];
    {
      const uint32_t t0[1] = {tile(buf * 2 + 0)};
      filtered_map<1, kTwo, kRbStride>(t0, lane_off, lane_off2, ta, tb, d);
#pragma unroll
      for (int ob = 0; ob < kOB; ++ob)
#pragma unroll
        for (int i = 0; i < 2; ++i) mx[ob][i] = __fmul2_rn(kx, make_float2(d[ob][2 * i], d[ob][2 * i + 1]));
      const uint32_t t1[1] = {tile(buf * 2 + 1)};
      filtered_map<1, kTwo, kRbStride>(t1, lane_off, lane_off2, ta, tb, d);
    }
#pragma unroll
    for (int ob = 0; ob < kOB; ++ob)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        my[ob][i] = __fmul2_rn(ky, make_float2(d[ob][2 * i], d[ob][2 * i + 1]));
        const float2 num = __ffma2_rn(two, __fmul2_rn(mx[ob][i], my[ob][i]), c6);
        const float2 den = __ffma2_rn(mx[ob][i], mx[ob][i], __ffma2_rn(my[ob][i], my[ob][i], c6));
        acc_a[ob][i] = __ffma2_rn(wc2, div2_approx(num, den), acc_a[ob][i]);
      }
    if (!producer) products_ready();
    {
      if (kTwo) {
        const uint32_t t2[2] = {tile(4), tile(5)};
        filtered_map<2, true, kRbStride>(t2, lane_off, lane_off2, ta, tb, d);
      } else {
        const uint32_t t2[1] = {tile(4 + 0 / 2)};
        filtered_map<1, false, kRbStride>(t2, lane_off, lane_off2, ta, tb, d);
      }
    }
#pragma unroll
    for (int ob = 0; ob < kOB; ++ob)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float2 m2 = __fmul2_rn(mx[ob][i], mx[ob][i]);
        const float2 vx = __ffma2_rn(kxx, make_float2(d[ob][2 * i], d[ob][2 * i + 1]), make_float2(-m2.x, -m2.y));
        vs[ob][i] = vx;
        float2 gd = __ffma2_rn(gk, mx[ob][i], ge);                         // >= ax^2 1e-12; kept out of the denormals
        gd = make_float2(fmaxf(gd.x, 1e-30f), fmaxf(gd.y, 1e-30f));
        acc_g[ob][i] = __fadd2_rn(acc_g[ob][i], div2_approx(vx, gd));
      }
    {
      if (kTwo) {
        const uint32_t t3[2] = {tile(6), tile(7)};
        filtered_map<2, true, kRbStride>(t3, lane_off, lane_off2, ta, tb, d);
      } else {
        const uint32_t t3[1] = {tile(4 + 2 / 2)};
        filtered_map<1, false, kRbStride>(t3, lane_off, lane_off2, ta, tb, d);
      }
    }
#pragma unroll
    for (int ob = 0; ob < kOB; ++ob)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float2 m2 = __fmul2_rn(my[ob][i], my[ob][i]);
        const float2 vy = __ffma2_rn(kyy, make_float2(d[ob][2 * i], d[ob][2 * i + 1]), make_float2(-m2.x, -m2.y));
        vs[ob][i] = __fadd2_rn(vs[ob][i], __fadd2_rn(vy, c6));
      }
    {
      if (kTwo) {
        const uint32_t t4[2] = {tile(8), tile(9)};
        filtered_map<2, true, kRbStride>(t4, lane_off, lane_off2, ta, tb, d);
      } else {
        const uint32_t t4[1] = {tile(4 + 4 / 2)};
        filtered_map<1, false, kRbStride>(t4, lane_off, lane_off2, ta, tb, d);
      }
    }
#pragma unroll
    for (int ob = 0; ob < kOB; ++ob)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float2 mm = __fmul2_rn(mx[ob][i], my[ob][i]);
        const float2 cov = __ffma2_rn(kxy, make_float2(d[ob][2 * i], d[ob][2 * i + 1]), make_float2(-mm.x, -mm.y));
        const float2 num = __ffma2_rn(two, cov, c6);
        acc_b[ob][i] = __ffma2_rn(wc2, div2_approx(num, vs[ob][i]), acc_b[ob][i]);
      }
  }

  if (!warp_live) return;
  const size_t mbase = ((size_t)b * csplit + split) * oh * ow;
#pragma unroll
  for (int ob = 0; ob < kOB; ++ob)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ox = ox0 + wx * 16 + (lane >> 2) + ((i & 2) ? 8 : 0);
      const int oy = oy0 + wy * 8 * kOB + ob * 8 + 2 * (lane & 3) + (i & 1);
      if (oy < oh && ox < ow) {
        const size_t o = mbase + (size_t)oy * ow + ox;
        amap[o] = (i & 1) ? acc_a[ob][i >> 1].y : acc_a[ob][i >> 1].x;
        bmap[o] = (i & 1) ? acc_b[ob][i >> 1].y : acc_b[ob][i >> 1].x;
        gmap[o] = (i & 1) ? acc_g[ob][i >> 1].y : acc_g[ob][i >> 1].x;
      }
    }
}

}  // namespace

bool window_mma_supported(int window, int h, int w, const void* fx, const void* fy) {
  // (w % 8 == 0 and 16-byte aligned planes: what a tensor map over the planes needs)
  return window >= 3 && window <= 21 && (window & 1) && h >= window && w >= window && (w % 8) == 0 &&
         (reinterpret_cast<uintptr_t>(fx) & 15) == 0 && (reinterpret_cast<uintptr_t>(fy) & 15) == 0;
}

static bool two_terms(int h, int w, long long one_term_min_pixels) {
  return one_term_min_pixels < 0 || (long long)h * w < one_term_min_pixels;
}

void window_mma_tiles(int h, int w, int window, int* tiles_x, int* tiles_y) {
  const int oh = h - window + 1, ow = w - window + 1;
  *tiles_x = ceil_div(ow, kTileOW);
  *tiles_y = ceil_div(oh, TcCfg<3, true>::kTileOH);   // (the same for the one-term kernel)
  static_assert(TcCfg<3, true>::kTileOH == TcCfg<3, false>::kTileOH, "tile height does not depend on the term count");
}

template <bool kTwo, bool kStrips>
static void launch_cfg(const CUtensorMap& mx, const CUtensorMap& my, int pairs, int c, int oh, int ow, int tx, int csplit, int choff,
                       const double* chan, const float* wt, const TcTaps& t, float* amap, float* bmap, float* gmap, cudaStream_t st) {
  using Cfg = TcCfg<3, kTwo>;
  ensure_dynamic_smem(window_mma_kernel<kTwo, 3, kStrips>, Cfg::kSmemBytes);
  dim3 grid(tx * ceil_div(oh, Cfg::kTileOH), csplit, pairs);
  window_mma_kernel<kTwo, 3, kStrips><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(mx, my, c, oh, ow, tx, csplit, choff, chan, wt, t,
                                                                                 amap, bmap, gmap);
}

int launch_window_mma(const bf16* fx, const bf16* fy, int pairs, int c, int h, int w, int window, int csplit, int choff,
                      const double* chan, const float* wt, const float* taps_f32, long long one_term_min_pixels, bool strips, float* amap,
                      float* bmap, float* gmap, cudaStream_t st) {
  TcTaps t;
  double sum = 0.0;
  for (int i = 0; i < 32; ++i) {
    t.g[i] = i < window ? __bfloat162float(__float2bfloat16_rn(taps_f32[i])) : 0.f;
    sum += (double)t.g[i];
  }
  t.inv_sum2 = (float)(1.0 / (sum * sum));
  const int oh = h - window + 1, ow = w - window + 1;
  // One bf16 term per operand or two.  The rounding of a one-term operand (2^-9 per element, ~4e-4 after the window
  // average) is an error of that fraction of E[f^2] in a variance; in the L2-normalised units of ADISTS.py:166-167 E[f^2]
  // is 1 / (h w) on average, so the error shrinks against the 1e-6 stabilisers c1 / c2 as the maps grow.  Measured on the
  // 1080p golden pair (profiles/r3o_window_one_term_sweep.log; distance of the score to the all-two-terms score, i.e.
  // the window arithmetic alone; the bf16 trunk itself sits 1.1e-5 from the fp32 reference): threshold 400k pixels
  // (levels 1-2 one-term) 3.6e-7, 100k (levels 1-3) 3.6e-6, 30k (levels 1-4) 5.0e-5, everything one-term 3.6e-4.  The
  // default, 100k, keeps the window rounding a factor 3 below the trunk's; smaller maps (deeper levels, small frames) keep
  // the exact two-term products and the hi + lo split between the passes.
  const bool two = two_terms(h, w, one_term_min_pixels);
  const int tx = ceil_div(ow, kTileOW);
  if (ceil_div(c, csplit) > kMaxCPer) return 0;
  CUtensorMap map_x, map_y;
  if (strips) {
    if (make_strip_map(&map_x, fx, pairs * c, h, w, 64, kInW / 8) != 0 || make_strip_map(&map_y, fy, pairs * c, h, w, 64, kInW / 8) != 0)
      return -1;   // strip planes have no other reader
  } else if (make_plane_map_sw64(&map_x, fx, pairs * c, h, w, TcCfg<3, true>::kInH) != 0 ||
             make_plane_map_sw64(&map_y, fy, pairs * c, h, w, TcCfg<3, true>::kInH) != 0) {
    return 0;   // the caller falls back to the CUDA-core window kernel
  }
  if (two) {
    if (strips) launch_cfg<true, true>(map_x, map_y, pairs, c, oh, ow, tx, csplit, choff, chan, wt, t, amap, bmap, gmap, st);
    else launch_cfg<true, false>(map_x, map_y, pairs, c, oh, ow, tx, csplit, choff, chan, wt, t, amap, bmap, gmap, st);
  } else {
    if (strips) launch_cfg<false, true>(map_x, map_y, pairs, c, oh, ow, tx, csplit, choff, chan, wt, t, amap, bmap, gmap, st);
    else launch_cfg<false, false>(map_x, map_y, pairs, c, oh, ow, tx, csplit, choff, chan, wt, t, amap, bmap, gmap, st);
  }
  return 1;
}

}  // namespace nqa
