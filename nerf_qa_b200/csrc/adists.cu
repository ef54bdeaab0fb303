// A-DISTS statistics stage (nerf_qa/ADISTS/ADISTS.py:71-100, 127-135, 147-197) on planar fp32 levels.
//
// Restructuring that preserves the reference's results:
//   * the 21x21 gaussian window is rank-1 -> two 21-tap passes through shared memory;
//   * windowed statistics are taken of the RAW features and scaled by the per-channel 1/||f||
//     afterwards (filtering is linear), which yields compute_prob's un-normalised mean / variance
//     from the same five filtered maps -> 5 filters per channel instead of 7;
//   * sum_c w_c((1-ps) T_c + ps S_c) = (1-ps) sum_c w_c T_c + ps sum_c w_c S_c, so the window kernel
//     reduces over channels on the fly and emits three 1-channel maps per level (sum w T, sum w S,
//     sum gamma_c); the [B,C,h,w] statistic maps never reach HBM;
//   * entropy weights depend only on the first image and are computed before the window pass;
//   * the coarse-to-fine ps chain runs on the small 1-channel maps with multi-block reductions whose
//     last block combines the partials in a fixed order (deterministic).
//
// Pass structure per micro-batch: (1) per-channel sums, (2) entropy, (3) channel weights,
// (4) window / global statistics per level, (5) ps chain + D per level (k = 5..0), (6) score.
#include <math.h>
#include <stdlib.h>

#include "adists.h"
#include "adists_tc.h"
#include "common.cuh"
#include "kernels.h"

namespace nqa {

namespace {

// feature planes are fp32 (level 0 = raw frames, forward_from_feats, fp32 trunk) or bf16 (bf16 trunk: half the bytes)
__device__ inline float ldf(const float* p) { return __ldg(p); }
__device__ inline float ldf(const bf16* p) { return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16); }
// Two adjacent plane elements as they come out of one load.  The window kernel keeps them RAW in registers while the
// loads are in flight and only unpacks when it writes the tile to shared memory one channel later (touching a loaded
// register earlier -- e.g. the bf16 -> fp32 shift -- would stall the warp on the load right away).
template <typename TI> struct Raw2;
template <> struct Raw2<float> {
  typedef float2 type;
  __device__ static inline type zero() { return make_float2(0.f, 0.f); }
  __device__ static inline type load(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  __device__ static inline type one(float a, float b) { return make_float2(a, b); }
  __device__ static inline void unpack(type t, float& a, float& b) { a = t.x; b = t.y; }
};
template <> struct Raw2<bf16> {
  typedef uint32_t type;
  __device__ static inline type zero() { return 0u; }
  __device__ static inline type load(const bf16* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
  __device__ static inline type one(float a, float b) {   // a, b are exact bf16 values
    return (__float_as_uint(a) >> 16) | (__float_as_uint(b) & 0xffff0000u);
  }
  __device__ static inline void unpack(type t, float& a, float& b) {
    a = __uint_as_float(t << 16); b = __uint_as_float(t & 0xffff0000u);
  }
};
__device__ inline void ld4(const float* p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ inline void ld4(const bf16* p, float (&v)[4]) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}

constexpr int WT = 32;           // window kernel: 32x32 output pixels per block
constexpr int kChan = 10;        // doubles per (pair, channel) in the chan table (slot 9: closed-form entropy)
constexpr int kPlaneTile = 256;  // pixels per tile of the fused NHWC -> planes + channel-sums kernel
constexpr int kRedBlocks = 512;  // blocks per frame pair in the map reductions (latency-bound gathers: fill the machine)

struct Taps {
  float g[kAdistsMaxWindow];
  float2 g2[kAdistsMaxWindow];   // (g, g): operand of the packed f32x2 FMAs of the window kernel
};

struct LvlScal {   // per (pair, level) scalars of the ps chain
  float mean, stdv, gmin, gmax;   // of gamma over the map
  float pmin, pmax;               // of ps_n * interp(prev)
  float d;                        // mean_hw((1-ps) A + ps B)
  float pad;
};

struct LevelPlan {
  int c, h, w, hw, choff;
  bool windowed;
  int oh, ow, on;          // window map (1x1 in the global-statistics fallback)
  int chunks;              // planar-sum chunks
  int tiles_x, tiles_y, csplit;
  int csplit_mma;          // channel split of the tensor-pipe window kernel (64 x 72 tiles: fewer blocks per map)
  int chunks8;             // pixel chunks of the fused NHWC -> planes + channel-sums kernel (levels with c % 64 == 0)
  size_t off_sum8;         // its partials [pairs][chunks8][8][c]
  int chunks_s;            // chunks of the strip-plane pooling + sums kernel (adists_strips.cu), 0 = level not served
  size_t off_sum_s;        // its partials [pairs][chunks_s][8][c]
  size_t off_sum, off_ent, off_a, off_b, off_g, off_ps;  // byte offsets into the workspace
};

struct Plan {
  LevelPlan lv[kNumLevels];
  size_t off_chan, off_e, off_wt, off_scal, off_red, off_cnt;
  size_t total;
};

size_t a256(size_t b) { return (b + 255) & ~(size_t)255; }

Plan make_plan(int pairs, int H, int W, int window) {
  Plan p;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = a256(off + bytes); return o; };
  int h = H, w = W;
  for (int k = 0; k < kNumLevels; ++k) {
    if (k >= 2) { h = pooled(h); w = pooled(w); }
    LevelPlan& L = p.lv[k];
    L.c = level_channels(k); L.h = h; L.w = w; L.hw = h * w; L.choff = level_offset(k);
    L.windowed = h >= window && w >= window;
    L.oh = L.windowed ? h - window + 1 : 1;
    L.ow = L.windowed ? w - window + 1 : 1;
    L.on = L.oh * L.ow;
    L.chunks = stats_chunks_planar(L.hw, L.c, pairs);
    L.tiles_x = ceil_div(L.ow, WT);
    L.tiles_y = ceil_div(L.oh, WT);
    L.csplit = 1;
    if (L.windowed) {
      // split the channel loop across blocks until the level fills ~2 waves (deep levels have few tiles)
      // (batch-independent on purpose: the channel split fixes the summation order of a frame's maps)
      const int blocks = L.tiles_x * L.tiles_y;
      int cs = ceil_div(2 * kNumSMs, blocks);
      if (cs > L.c / 8) cs = L.c / 8;
      if (cs < 1) cs = 1;
      L.csplit = cs;
    }
    L.csplit_mma = 1;
    if (L.windowed) {
      int tx, ty;
      window_mma_tiles(L.h, L.w, window, &tx, &ty);
      int cs = ceil_div(2 * kNumSMs, tx * ty);   // one 12-warp block per SM
      if (cs > L.c / 8) cs = L.c / 8;
      if (cs < 1) cs = 1;
      L.csplit_mma = cs;
    }
    L.chunks8 = 0;
    L.off_sum8 = 0;
    // (maps of fewer than 256 pixels keep the separate kernels: their entropies are tiny or exactly zero -- 1x1 maps --
    // and the per-level normalisation of ADISTS.py:134 would amplify the rounding of the closed form)
    if (L.c % 64 == 0 && L.hw >= kPlaneTile) {
      // ~512 blocks per frame pair whatever the level (batch-independent: the chunking fixes the summation order)
      const int ntiles = ceil_div(L.hw, kPlaneTile);
      int ch8 = 512 / (L.c / 64);
      if (ch8 > ntiles) ch8 = ntiles;
      L.chunks8 = ch8 < 1 ? 1 : ch8;
      L.off_sum8 = take((size_t)pairs * L.chunks8 * 8 * L.c * sizeof(float));
    }
    L.chunks_s = 0;
    L.off_sum_s = 0;
    if (k >= 1 && L.c % 64 == 0 && (w & 7) == 0 && L.windowed) {
      int spb, n_sg;
      adists_strips_plan(L.c, h, w, &spb, &n_sg, &L.chunks_s);
      L.off_sum_s = take((size_t)pairs * L.chunks_s * 8 * L.c * sizeof(float));
    }
    const int cs_max = L.csplit > L.csplit_mma ? L.csplit : L.csplit_mma;
    L.off_sum = take((size_t)pairs * L.chunks * 7 * L.c * sizeof(float));
    L.off_ent = take((size_t)pairs * L.chunks * L.c * sizeof(float));
    L.off_a = take((size_t)pairs * cs_max * L.on * sizeof(float));
    L.off_b = take((size_t)pairs * cs_max * L.on * sizeof(float));
    L.off_g = take((size_t)pairs * cs_max * L.on * sizeof(float));
    L.off_ps = take((size_t)pairs * L.on * sizeof(float));
  }
  p.off_chan = take((size_t)pairs * kTotalChannels * kChan * sizeof(double));
  p.off_e = take((size_t)pairs * kTotalChannels * sizeof(double));
  p.off_wt = take((size_t)pairs * kTotalChannels * sizeof(float));
  p.off_scal = take((size_t)pairs * kNumLevels * sizeof(LvlScal));
  p.off_red = take((size_t)pairs * kRedBlocks * 4 * sizeof(double));
  p.off_cnt = take((size_t)pairs * sizeof(unsigned int));
  p.total = off + 256;
  return p;
}

// ================================================================================================
// (1) per-(pair, channel) sums over the map: x, y, x^2, y^2, xy, relu(x), relu(x)^2
// grid = (chunks, C, pairs); partial[((b*chunks + chunk)*7 + s)*C + ch]
// ================================================================================================
template <typename TI>
__global__ void __launch_bounds__(256)
sums7_kernel(const TI* __restrict__ x, const TI* __restrict__ y, int c, int hw, int chunks,
             float* __restrict__ partials) {
  const int chunk = blockIdx.x, ch = blockIdx.y, b = blockIdx.z;
  const int per = ceil_div(ceil_div(hw, chunks), 4) * 4;
  const int p0 = chunk * per, p1 = min(hw, p0 + per);
  const TI* px = x + ((size_t)b * c + ch) * hw;
  const TI* py = y + ((size_t)b * c + ch) * hw;
  constexpr uintptr_t kAlign = 4 * sizeof(TI) - 1;
  float s[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto acc = [&](float a, float d) {
    const float r = fmaxf(a, 0.f);
    s[0] += a; s[1] += d;
    s[2] = fmaf(a, a, s[2]); s[3] = fmaf(d, d, s[3]); s[4] = fmaf(a, d, s[4]);
    s[5] += r; s[6] = fmaf(r, r, s[6]);
  };
  const bool vec_ok = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(px) & kAlign) == 0) &&
                      ((reinterpret_cast<uintptr_t>(py) & kAlign) == 0);
  if (vec_ok && p1 > p0) {
    const int n4 = (p1 - p0) / 4;
#pragma unroll 2
    for (int i = threadIdx.x; i < n4; i += 256) {
      float a[4], d[4];
      ld4(px + p0 + 4 * i, a);
      ld4(py + p0 + 4 * i, d);
      acc(a[0], d[0]); acc(a[1], d[1]); acc(a[2], d[2]); acc(a[3], d[3]);
    }
    for (int p = p0 + n4 * 4 + threadIdx.x; p < p1; p += 256) acc(ldf(px + p), ldf(py + p));
  } else {
    for (int p = p0 + threadIdx.x; p < p1; p += 256) acc(ldf(px + p), ldf(py + p));
  }
  __shared__ float sm[7][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const float v = warp_sum(s[i]);
    if (lane == 0) sm[i][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += sm[threadIdx.x][q];
    partials[(((size_t)b * chunks + chunk) * 7 + threadIdx.x) * c + ch] = v;
  }
}

// ================================================================================================
// (1') bf16 trunk: the NHWC level -> bf16 planes transposition fused with the channel sums, so that the
// planes are not read again for sums and entropy.  Eight sums per (pair, channel): x, y, x^2, y^2, xy, relu(x),
// relu(x)^2 and relu(x) log2 relu(x) -- the last gives the entropy of ADISTS.py:127-135 in closed form:
//   p = r / Z,  Z = sum r + ||r|| 1e-12,   E = -sum p log2 p = (log2(Z) sum r - sum r log2 r) / Z
// (the reference's log2(p + 1e-12) differs only where p <~ 1e-12, i.e. by < 1e-10 of an entropy of ~20 bits).
// Block = 256 pixels x 64 channels of the x AND the y image of one pair per tile, `per` tiles per block;
// sums accumulate per thread in the load-phase registers (a thread owns 8 channels of 8 pixels per tile) and are
// combined at the end in a fixed order.  grid = (chunks8, c/64, pairs), block = 256, dynamic smem 2 x 33 KB.
// partial[((b*chunks8 + chunk)*8 + s)*C + ch]
// ================================================================================================
__global__ void __launch_bounds__(256)
planes_sums_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int pairs, int c, int hw, int chunks8,
                   float* __restrict__ partials, int strip_w, int strip_h) {
  extern __shared__ uint32_t ps_smem[];
  uint32_t (*tx)[33] = reinterpret_cast<uint32_t (*)[33]>(ps_smem);                      // [pixel][channel pair]
  uint32_t (*ty)[33] = reinterpret_cast<uint32_t (*)[33]>(ps_smem + kPlaneTile * 33);
  const int chunk = blockIdx.x, c0 = blockIdx.y * 64, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lp = lane >> 3, lc = lane & 7;
  const int ntiles = ceil_div(hw, kPlaneTile), per = ceil_div(ntiles, chunks8);
  const int t0 = chunk * per, t1 = min(ntiles, t0 + per);
  const bf16* srcx = in + (size_t)b * hw * c + c0;
  const bf16* srcy = in + (size_t)(pairs + b) * hw * c + c0;
  bf16* dstx = out + (size_t)b * c * hw;
  bf16* dsty = out + (size_t)(pairs + b) * c * hw;
  const bool even = (hw & 1) == 0;     // 4-byte stores need even plane sizes (plane starts stay 4-byte aligned)

  // acc[e][s]: sums of the channel pair (lc*8 + 2e, lc*8 + 2e + 1) -- the two bf16 halves of a loaded word -- on the
  // packed fp32 pipe
  float2 acc[4][8];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[e][q] = make_float2(0.f, 0.f);

  for (int t = t0; t < t1; ++t) {
    const int p0 = t * kPlaneTile;
#pragma unroll 2
    for (int k = 0; k < 8; ++k) {
      const int pix = (k * 8 + warp) * 4 + lp, p = p0 + pix;
      const size_t off = (size_t)min(p, hw - 1) * c;   // clamped, unconditional (masked below)
      const uint4 vx = __ldg(reinterpret_cast<const uint4*>(srcx + off) + lc);
      const uint4 vy = __ldg(reinterpret_cast<const uint4*>(srcy + off) + lc);
      uint32_t* rx = &tx[pix][lc * 4];
      uint32_t* ry = &ty[pix][lc * 4];
      rx[0] = vx.x; rx[1] = vx.y; rx[2] = vx.z; rx[3] = vx.w;
      ry[0] = vy.x; ry[1] = vy.y; ry[2] = vy.z; ry[3] = vy.w;
      if (p < hw) {
        const uint32_t wx[4] = {vx.x, vx.y, vx.z, vx.w}, wy[4] = {vy.x, vy.y, vy.z, vy.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = make_float2(__uint_as_float(wx[e] << 16), __uint_as_float(wx[e] & 0xffff0000u));
          const float2 d = make_float2(__uint_as_float(wy[e] << 16), __uint_as_float(wy[e] & 0xffff0000u));
          const float2 r = make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
          // r log2 r with the logarithm of max(r, 1e-30): 0 * (-99.7) = 0 for the rectified zeros, no select
          const float2 lg = make_float2(__log2f(fmaxf(a.x, 1e-30f)), __log2f(fmaxf(a.y, 1e-30f)));
          float2* s = acc[e];
          s[0] = __fadd2_rn(s[0], a); s[1] = __fadd2_rn(s[1], d);
          s[2] = __ffma2_rn(a, a, s[2]); s[3] = __ffma2_rn(d, d, s[3]); s[4] = __ffma2_rn(a, d, s[4]);
          s[5] = __fadd2_rn(s[5], r); s[6] = __ffma2_rn(r, r, s[6]);
          s[7] = __ffma2_rn(r, lg, s[7]);
        }
      }
    }
    __syncthreads();
    for (int ch = warp; ch < 64; ch += 8) {
      unsigned short* rowx = reinterpret_cast<unsigned short*>(dstx + (size_t)(c0 + ch) * hw);
      unsigned short* rowy = reinterpret_cast<unsigned short*>(dsty + (size_t)(c0 + ch) * hw);
#pragma unroll
      for (int q = 0; q < kPlaneTile / 64; ++q) {
        const int pp = q * 32 + lane, p = p0 + 2 * pp;
        const uint32_t ax = tx[2 * pp][ch >> 1], bx = tx[2 * pp + 1][ch >> 1];
        const uint32_t ay = ty[2 * pp][ch >> 1], by = ty[2 * pp + 1][ch >> 1];
        const uint32_t lox = (ch & 1) ? (ax >> 16) : (ax & 0xffffu), hix = (ch & 1) ? (bx >> 16) : (bx & 0xffffu);
        const uint32_t loy = (ch & 1) ? (ay >> 16) : (ay & 0xffffu), hiy = (ch & 1) ? (by >> 16) : (by & 0xffffu);
        if (strip_w > 0) {   // strip layout [w / 8][h][8]: the pixel pair (p, p + 1) stays inside one strip row (w % 8 == 0)
          if (p < hw) {
            const int yy = p / strip_w, xx = p - yy * strip_w;
            const size_t o = ((size_t)(xx >> 3) * strip_h + yy) * 8 + (xx & 7);
            *reinterpret_cast<uint32_t*>(rowx + o) = lox | (hix << 16);
            *reinterpret_cast<uint32_t*>(rowy + o) = loy | (hiy << 16);
          }
        } else if (even && p + 1 < hw) {
          *reinterpret_cast<uint32_t*>(rowx + p) = lox | (hix << 16);
          *reinterpret_cast<uint32_t*>(rowy + p) = loy | (hiy << 16);
        } else {
          if (p < hw) { rowx[p] = (unsigned short)lox; rowy[p] = (unsigned short)loy; }
          if (p + 1 < hw) { rowx[p + 1] = (unsigned short)hix; rowy[p + 1] = (unsigned short)hiy; }
        }
      }
    }
    __syncthreads();
  }
  // combine: the four pixel lanes (lp) of a channel group by shuffles, the eight warps through shared memory, in order
  float* red = reinterpret_cast<float*>(ps_smem);   // [warp][lc][j][s] = 8 * 8 * 64 floats
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v = (j & 1) ? acc[j >> 1][q].y : acc[j >> 1][q].x;
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lp == 0) red[((warp * 8 + lc) * 8 + j) * 8 + q] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += 256) {   // i = (channel within the group) * 8 + sum
    const int chl = i >> 3, q = i & 7;             // chl = lc * 8 + j
    float v = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) v += red[(wq * 64 + chl) * 8 + q];
    partials[(((size_t)b * chunks8 + chunk) * 8 + q) * c + c0 + chl] = v;
  }
}

// chan[b][1475][kChan] (double): 1/max(||x||,eps), 1/max(||y||,eps), max(||relu x||,eps), mean x, mean y, E[x^2], E[y^2],
// E[xy], sum_hw(relu x / max(||relu x||,eps))
struct SumLevels {
  const float* partials[kNumLevels];
  int chunks[kNumLevels];
  int hw[kNumLevels];
  int fused[kNumLevels];   // 1: eight sums per chunk from planes_sums_kernel (entropy in closed form), 0: seven
};

// grid = (ceil(1475 / 32), pairs), block = 256: eight threads per channel split the chunks (thread t takes chunks t,
// t + 8, ...), combined in lane order -- a fixed summation order.
__global__ void __launch_bounds__(256)
chan_kernel(SumLevels sl, double* __restrict__ chan) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x >> 3), part = threadIdx.x & 7;
  const bool live = c < kTotalChannels;
  const int cl = live ? c : kTotalChannels - 1;
  const int k = cl < 3 ? 0 : cl < 67 ? 1 : cl < 195 ? 2 : cl < 451 ? 3 : cl < 963 ? 4 : 5;
  const int C = level_channels(k), cc = cl - level_offset(k), chunks = sl.chunks[k];
  const int ns = sl.fused[k] ? 8 : 7;
  const float* base = sl.partials[k] + (size_t)b * chunks * ns * C + cc;
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int ch = part; ch < chunks; ch += 8)
    for (int i = 0; i < ns; ++i) s[i] += (double)base[((size_t)ch * ns + i) * C];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s[i] += __shfl_xor_sync(0xffffffffu, s[i], 1);
    s[i] += __shfl_xor_sync(0xffffffffu, s[i], 2);
    s[i] += __shfl_xor_sync(0xffffffffu, s[i], 4);
  }
  if (!live || part != 0) return;
  const double n = (double)sl.hw[k];
  const double eps = 1e-12;
  const double nx = fmax(sqrt(s[2]), eps), ny = fmax(sqrt(s[3]), eps), nr = fmax(sqrt(s[6]), eps);
  double* o = chan + ((size_t)b * kTotalChannels + c) * kChan;
  o[0] = 1.0 / nx;
  o[1] = 1.0 / ny;
  o[2] = nr;                     // p = (relu(x) / o[2]) / (o[8] + 1e-12)  (ADISTS.py:130-132), formed in fp32 like
  o[8] = s[5] / nr;              // the reference so that degenerate 1x1 maps give exactly p = 1 and zero entropy
  o[3] = s[0] / n;
  o[4] = s[1] / n;
  o[5] = s[2] / n;
  o[6] = s[3] / n;
  o[7] = s[4] / n;
  if (sl.fused[k]) {             // entropy in closed form (see planes_sums_kernel)
    const double z = s[5] + nr * 1e-12;
    o[9] = s[5] > 0.0 ? (log2(z) * s[5] - s[7]) / z : 0.0;
  } else {
    o[9] = 0.0;
  }
}

// ================================================================================================
// (2) entropy of the first image's channels: E_c = -sum_hw p log2(p + 1e-12)   (ADISTS.py:133)
// grid = (chunks, C, pairs); partial[(b*chunks + chunk)*C + ch]
// ================================================================================================
template <typename TI>
__global__ void __launch_bounds__(256)
entropy_kernel(const TI* __restrict__ x, int c, int hw, int chunks, int choff, const double* __restrict__ chan,
               float* __restrict__ partials) {
  const int chunk = blockIdx.x, ch = blockIdx.y, b = blockIdx.z;
  const int per = ceil_div(ceil_div(hw, chunks), 4) * 4;
  const int p0 = chunk * per, p1 = min(hw, p0 + per);
  const TI* px = x + ((size_t)b * c + ch) * hw;
  const double* cs = chan + ((size_t)b * kTotalChannels + choff + ch) * kChan;
  const float nr = (float)cs[2], den = (float)cs[8] + 1e-12f;
  const float scale = (float)(1.0 / ((double)nr * (double)den));
  const bool exact = hw == 1;   // a 1x1 map: the reference's (relu(x)/||relu(x)||)/(sum + 1e-12) is exactly 1 (or 0), its
                                // entropy exactly 0, and the channel normalisation that follows turns any rounding into O(1)
  float s = 0.f;
  auto acc = [&](float a) {
    const float r = fmaxf(a, 0.f);
    const float p = exact ? (r / nr) / den : r * scale;
    // bf16 planes (2e-3 tolerance): the MUFU logarithm (2^-22 relative away from 1, tiny absolute error near 1, where
    // p log p vanishes anyway) instead of the 20-instruction libm path this kernel was bound by
    s -= p * (sizeof(TI) == 2 ? __log2f(p + 1e-12f) : log2f(p + 1e-12f));
  };
  const bool vec_ok = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(px) & (4 * sizeof(TI) - 1)) == 0);
  if (vec_ok && p1 > p0) {
    const int n4 = (p1 - p0) / 4;
#pragma unroll 2
    for (int i = threadIdx.x; i < n4; i += 256) {
      float a[4];
      ld4(px + p0 + 4 * i, a);
      acc(a[0]); acc(a[1]); acc(a[2]); acc(a[3]);
    }
    for (int p = p0 + n4 * 4 + threadIdx.x; p < p1; p += 256) acc(ldf(px + p));
  } else {
    for (int p = p0 + threadIdx.x; p < p1; p += 256) acc(ldf(px + p));
  }
  __shared__ float sm[33];
  const float tot = block_sum<float>(s, sm);
  if (threadIdx.x == 0) partials[((size_t)b * chunks + chunk) * c + ch] = tot;
}

// ================================================================================================
// (3) channel weights: per-level normalise * C_k, concat, normalise, clamp(mean +- 0.5 std),
// renormalise (ADISTS.py:134-135, 155-160).  grid = pairs.
// ================================================================================================
struct EntLevels {
  const float* partials[kNumLevels];
  int chunks[kNumLevels];   // 0: the level's entropies are already in chan[..][9] (closed form)
};

__global__ void __launch_bounds__(256)
weights_kernel(EntLevels el, const double* __restrict__ chan, float* __restrict__ wt) {
  __shared__ double e[kTotalChannels];
  __shared__ double red[33];
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < kTotalChannels; c += 256) {
    const int k = c < 3 ? 0 : c < 67 ? 1 : c < 195 ? 2 : c < 451 ? 3 : c < 963 ? 4 : 5;
    const int C = level_channels(k), cc = c - level_offset(k), chunks = el.chunks[k];
    double s = 0;
    if (chunks == 0) {
      s = chan[((size_t)b * kTotalChannels + c) * kChan + 9];
    } else {
      const float* base = el.partials[k] + (size_t)b * chunks * C + cc;
      for (int ch = 0; ch < chunks; ++ch) s += (double)base[(size_t)ch * C];
    }
    e[c] = s;
  }
  __syncthreads();
  for (int k = 0; k < kNumLevels; ++k) {
    const int C = level_channels(k), off = level_offset(k);
    double part = 0;
    for (int c = threadIdx.x; c < C; c += 256) part += e[off + c];
    const double tot = block_sum<double>(part, red);
    for (int c = threadIdx.x; c < C; c += 256) e[off + c] = e[off + c] / (tot + 1e-12) * (double)C;
    __syncthreads();
  }
  double part = 0;
  for (int c = threadIdx.x; c < kTotalChannels; c += 256) part += e[c];
  const double tot = block_sum<double>(part, red);
  for (int c = threadIdx.x; c < kTotalChannels; c += 256) e[c] /= tot;
  __syncthreads();
  part = 0;
  for (int c = threadIdx.x; c < kTotalChannels; c += 256) part += e[c];
  const double mean = block_sum<double>(part, red) / kTotalChannels;
  part = 0;
  for (int c = threadIdx.x; c < kTotalChannels; c += 256) part += (e[c] - mean) * (e[c] - mean);
  const double sd = sqrt(block_sum<double>(part, red) / kTotalChannels);
  const double lo = mean - 0.5 * sd, hi = mean + 0.5 * sd;
  part = 0;
  for (int c = threadIdx.x; c < kTotalChannels; c += 256) {
    e[c] = fmin(fmax(e[c], lo), hi);
    part += e[c];
  }
  const double tot2 = block_sum<double>(part, red);
  for (int c = threadIdx.x; c < kTotalChannels; c += 256) wt[(size_t)b * kTotalChannels + c] = (float)(e[c] / tot2);
}

// ================================================================================================
// (4a) window statistics.  Block = 32x32 output pixels of one frame pair, loops over a channel range:
// loads the (32+WIN-1)^2 raw tiles of x and y, horizontal 5-quantity pass into shared memory, vertical
// pass in registers, then per-pixel T, S, gamma accumulated over channels.
// grid = (tiles_x*tiles_y, csplit, pairs); out maps [pairs][csplit][oh*ow].
// ================================================================================================
template <int WIN, typename TI>
__global__ void __launch_bounds__(256, 2)
window_kernel(const TI* __restrict__ fx, const TI* __restrict__ fy, int c, int h, int w, int oh, int ow,
              int tiles_x, int csplit, int choff, const double* __restrict__ chan, const float* __restrict__ wt,
              Taps taps, float* __restrict__ amap, float* __restrict__ bmap, float* __restrict__ gmap) {
  constexpr int IN = WT + WIN - 1;
  constexpr int LDI = IN | 1;     // odd row stride: conflict-free when lanes walk down rows
  constexpr int LDH = WT + 1;
  constexpr int RX = 8;           // output columns per thread in the horizontal pass
  constexpr int RY = 4;           // output rows per thread in the vertical pass
  static_assert(IN <= 64, "horizontal pass maps rows onto two warps per column group");
  extern __shared__ float smem[];
  float* sx = smem;
  float* sy = sx + IN * LDI;
  float2* hb01 = reinterpret_cast<float2*>(sy + IN * LDI + ((2 * IN * LDI) & 1));   // [IN][LDH] (G_h x, G_h y), 8-byte aligned
  float2* hb23 = hb01 + IN * LDH;                                                  // [IN][LDH] (G_h x^2, G_h y^2)
  float* hb4 = reinterpret_cast<float*>(hb23 + IN * LDH);                          // [IN][LDH] G_h xy

  const int tile = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
  const int oy0 = (tile / tiles_x) * WT, ox0 = (tile % tiles_x) * WT;
  const int c_per = ceil_div(c, csplit);
  const int c_begin = split * c_per, c_end = min(c, c_begin + c_per);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // horizontal role: column group jg (8 output columns), input row hr
  const int jg = warp & 3, hr = (warp >> 2) * 32 + lane;
  // vertical role: output column vj, output rows vi0..vi0+3
  const int vj = lane, vi0 = warp * RY;

  float acc_a[RY], acc_b[RY], acc_g[RY];
#pragma unroll
  for (int i = 0; i < RY; ++i) acc_a[i] = acc_b[i] = acc_g[i] = 0.f;

  // software pipeline over channels: the global loads of channel ch+1 are issued right after channel ch's tile
  // has been written to shared memory and stay in flight during its two filter passes
  // the tile is fetched as element PAIRS (one 4- or 8-byte load per two pixels) when the rows allow it (even width =>
  // even global column, aligned planes); IN is even, so a pair never straddles a tile row
  static_assert(IN % 2 == 0, "pair loads assume an even tile width");
  constexpr int NPAIR = IN * IN / 2;
  constexpr int NLD = (NPAIR + 255) / 256;
  const bool pair_ok = (w & 1) == 0 && (reinterpret_cast<uintptr_t>(fx) & (2 * sizeof(TI) - 1)) == 0 &&
                       (reinterpret_cast<uintptr_t>(fy) & (2 * sizeof(TI) - 1)) == 0;
  typedef Raw2<TI> R2;
  typename R2::type pre_x[NLD], pre_y[NLD];
  // Pair path: UNCONDITIONAL loads from a clamped address (a load under a branch makes the compiler wait for it at the
  // reconvergence point, which serialises the NLD global-load latencies of a channel); the out-of-image elements are
  // zeroed when the tile is written to shared memory.  Offsets and validity do not depend on the channel.
  uint32_t pre_valid = 0;
  int pre_off[NLD];
#pragma unroll
  for (int q = 0; q < NLD; ++q) {
    const int idx = threadIdx.x + q * 256;
    const int r = idx / (IN / 2), col = 2 * (idx - r * (IN / 2));
    const int gy = oy0 + r, gx = ox0 + col;
    if (idx < NPAIR && gy < h && gx < w) pre_valid |= 1u << q;
    pre_off[q] = min(gy, h - 1) * w + min(gx, max(w - 2, 0));
  }
  auto fetch = [&](int ch) {
    const TI* px = fx + ((size_t)b * c + ch) * h * w;
    const TI* py = fy + ((size_t)b * c + ch) * h * w;
    if (pair_ok) {
#pragma unroll
      for (int q = 0; q < NLD; ++q) {
        pre_x[q] = R2::load(px + pre_off[q]);
        pre_y[q] = R2::load(py + pre_off[q]);
      }
      return;
    }
#pragma unroll
    for (int q = 0; q < NLD; ++q) {                 // odd width: scalar loads (rare; latency is exposed here)
      const int idx = threadIdx.x + q * 256;
      const int r = idx / (IN / 2), col = 2 * (idx - r * (IN / 2));
      const int gy = oy0 + r, gx = ox0 + col;
      pre_x[q] = pre_y[q] = R2::zero();
      if (idx < NPAIR && gy < h && gx < w) {
        const size_t o = (size_t)gy * w + gx;
        const bool two = gx + 1 < w;
        pre_x[q] = R2::one(ldf(px + o), two ? ldf(px + o + 1) : 0.f);
        pre_y[q] = R2::one(ldf(py + o), two ? ldf(py + o + 1) : 0.f);
      }
    }
  };
  if (c_begin < c_end) fetch(c_begin);
  for (int ch = c_begin; ch < c_end; ++ch) {
    __syncthreads();  // previous channel's vertical pass has finished reading hb; sx/sy are free
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      const int idx = threadIdx.x + q * 256;
      if (idx < NPAIR) {
        const int r = idx / (IN / 2), col = 2 * (idx - r * (IN / 2));
        float a0, a1, d0, d1;
        R2::unpack(pre_x[q], a0, a1);
        R2::unpack(pre_y[q], d0, d1);
        if (!((pre_valid >> q) & 1)) a0 = a1 = d0 = d1 = 0.f;
        sx[r * LDI + col] = a0;
        sx[r * LDI + col + 1] = a1;
        sy[r * LDI + col] = d0;
        sy[r * LDI + col + 1] = d1;
      }
    }
    __syncthreads();
    if (ch + 1 < c_end) fetch(ch + 1);
    // The five filtered planes are carried as two packed pairs (x, y), (x^2, y^2) and one scalar (x y): the sm_100
    // f32x2 FMAs halve the issue slots of the 21-tap passes (the kernel is issue-bound), same IEEE arithmetic per lane.
    if (hr < IN) {
      float2 o01[RX], o23[RX];
      float o4[RX];
#pragma unroll
      for (int j = 0; j < RX; ++j) { o01[j] = o23[j] = make_float2(0.f, 0.f); o4[j] = 0.f; }
      const float* rx = sx + hr * LDI + jg * RX;
      const float* ry = sy + hr * LDI + jg * RX;
#pragma unroll
      for (int t = 0; t < RX + WIN - 1; ++t) {
        const float2 ad = make_float2(rx[t], ry[t]);
        const float2 sq = __fmul2_rn(ad, ad);
        const float cr = ad.x * ad.y;
#pragma unroll
        for (int j = 0; j < RX; ++j) {
          if (t - j >= 0 && t - j < WIN) {
            o01[j] = __ffma2_rn(taps.g2[t - j], ad, o01[j]);
            o23[j] = __ffma2_rn(taps.g2[t - j], sq, o23[j]);
            o4[j] = fmaf(taps.g[t - j], cr, o4[j]);
          }
        }
      }
      float2* h01 = hb01 + hr * LDH + jg * RX;
      float2* h23 = hb23 + hr * LDH + jg * RX;
      float* h4 = hb4 + hr * LDH + jg * RX;
#pragma unroll
      for (int j = 0; j < RX; ++j) {
        h01[j] = o01[j];
        h23[j] = o23[j];
        h4[j] = o4[j];
      }
    }
    __syncthreads();
    float2 v01[RY], v23[RY];
    float v4[RY];
#pragma unroll
    for (int i = 0; i < RY; ++i) { v01[i] = v23[i] = make_float2(0.f, 0.f); v4[i] = 0.f; }
#pragma unroll
    for (int t = 0; t < RY + WIN - 1; ++t) {
      const float2 a01 = hb01[(vi0 + t) * LDH + vj];
      const float2 a23 = hb23[(vi0 + t) * LDH + vj];
      const float a4 = hb4[(vi0 + t) * LDH + vj];
#pragma unroll
      for (int i = 0; i < RY; ++i)
        if (t - i >= 0 && t - i < WIN) {
          v01[i] = __ffma2_rn(taps.g2[t - i], a01, v01[i]);
          v23[i] = __ffma2_rn(taps.g2[t - i], a23, v23[i]);
          v4[i] = fmaf(taps.g[t - i], a4, v4[i]);
        }
    }
    const double* cs = chan + ((size_t)b * kTotalChannels + choff + ch) * kChan;
    const float ax = (float)cs[0], ay = (float)cs[1];
    const float wc = wt[(size_t)b * kTotalChannels + choff + ch];
#pragma unroll
    for (int i = 0; i < RY; ++i) {
      const float mxr = v01[i].x, myr = v01[i].y;
      const float mx = ax * mxr, my = ay * myr;
      const float vx = ax * ax * v23[i].x - mx * mx;
      const float vy = ay * ay * v23[i].y - my * my;
      const float cov = ax * ay * v4[i] - mx * my;
      const float t_ = (2.f * mx * my + 1e-6f) / (mx * mx + my * my + 1e-6f);
      const float s_ = (2.f * cov + 1e-6f) / (vx + vy + 1e-6f);
      acc_a[i] = fmaf(wc, t_, acc_a[i]);
      acc_b[i] = fmaf(wc, s_, acc_b[i]);
      acc_g[i] += (v23[i].x - mxr * mxr) / (mxr + 1e-12f);   // compute_prob, raw features (ADISTS.py:81-84)
    }
  }
  const size_t mbase = ((size_t)b * csplit + split) * oh * ow;
#pragma unroll
  for (int i = 0; i < RY; ++i) {
    const int oy = oy0 + vi0 + i, ox = ox0 + vj;
    if (oy < oh && ox < ow) {
      const size_t o = mbase + (size_t)oy * ow + ox;
      amap[o] = acc_a[i];
      bmap[o] = acc_b[i];
      gmap[o] = acc_g[i];
    }
  }
}

// (4b) global-statistics fallback for a level smaller than the window (ADISTS.py:91-97, 175-180).
// grid = pairs; writes the 1x1 "maps".
__global__ void __launch_bounds__(256)
global_level_kernel(int c, int choff, const double* __restrict__ chan, const float* __restrict__ wt,
                    float* __restrict__ amap, float* __restrict__ bmap, float* __restrict__ gmap) {
  __shared__ double red[33];
  const int b = blockIdx.x;
  double pa = 0, pb = 0, pg = 0;
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    const double* cs = chan + ((size_t)b * kTotalChannels + choff + ch) * kChan;
    const double ax = cs[0], ay = cs[1];
    const double mxr = cs[3], myr = cs[4];
    const double mx = ax * mxr, my = ay * myr;
    const double vx = ax * ax * cs[5] - mx * mx, vy = ay * ay * cs[6] - my * my;
    const double cov = ax * ay * cs[7] - mx * my;
    const double t_ = (2.0 * mx * my + 1e-6) / (mx * mx + my * my + 1e-6);
    const double s_ = (2.0 * cov + 1e-6) / (vx + vy + 1e-6);
    const double wc = (double)wt[(size_t)b * kTotalChannels + choff + ch];
    pa += wc * t_;
    pb += wc * s_;
    pg += (cs[5] - mxr * mxr) / (mxr + 1e-12);
  }
  const double ta = block_sum<double>(pa, red);
  const double tb = block_sum<double>(pb, red);
  const double tg = block_sum<double>(pg, red);
  if (threadIdx.x == 0) {
    amap[b] = (float)ta;
    bmap[b] = (float)tb;
    gmap[b] = (float)tg;
  }
}

// ================================================================================================
// (5) ps chain on the 1-channel maps (ADISTS.py:85-99).  Three multi-block reductions per windowed
// level; the last block of each frame pair combines the block partials in index order.
// ================================================================================================
struct MapArgs {
  const float* amap; const float* bmap; float* gmap;   // [pairs][csplit][on]
  int csplit, stride, on, oh, ow, c;   // slices summed per pixel (1 after reduce_splits_kernel), slices per pair
  const float* prev;     // coarser ps_prod [pairs][ph*pw], or nullptr = ones
  int ph, pw;
  float* ps_out;         // [pairs][on]
  LvlScal* scal;         // [pairs][6]
  int level;
  double* red;           // [pairs][kRedBlocks][4]
  unsigned int* cnt;     // [pairs]
  int store_d;           // as_map: keep the per-pixel D value in slot 0 of gmap
};

__device__ inline float map_sum(const float* m, int stride, int on, int b, int i) { return m[(size_t)b * stride * on + i]; }

// bilinear, align_corners=True (torch upsample_bilinear2d semantics); sy / sx are the per-map scale factors
// (in - 1) / (out - 1), computed once per thread
__device__ inline float interp_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }
__device__ inline float interp_prev(const float* prev, int ph, int pw, int b, int oy, int ox, float sy, float sx) {
  if (!prev) return 1.f;
  const float* p = prev + (size_t)b * ph * pw;
  const float fy = sy * oy, fxx = sx * ox;
  const int y0 = (int)fy, x0 = (int)fxx;
  const int y1 = y0 + (y0 < ph - 1 ? 1 : 0), x1 = x0 + (x0 < pw - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fxx - x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  return hy * (hx * p[y0 * pw + x0] + lx * p[y0 * pw + x1]) + ly * (hx * p[y1 * pw + x0] + lx * p[y1 * pw + x1]);
}

__device__ inline float sigmoidf_ref(float z) { return 1.f / (1.f + expf(-z)); }

// returns true in the last block of this pair to finish (after all partials are visible)
__device__ inline bool publish_partials(double* red, unsigned int* cnt, int b, const double (&vals)[4]) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    double* r = red + ((size_t)b * kRedBlocks + blockIdx.x) * 4;
    r[0] = vals[0]; r[1] = vals[1]; r[2] = vals[2]; r[3] = vals[3];
    __threadfence();
    const unsigned int done = atomicAdd(cnt + b, 1u);
    is_last = done == gridDim.x - 1;
    if (is_last) cnt[b] = 0;  // re-arm for the next reduction (stream-ordered)
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// Sum of the channel-split slices of the three maps into slice 0, in slice order (deterministic); afterwards the ps
// chain reads one slice per map instead of csplit strided ones in each of its three kernels.
__global__ void __launch_bounds__(256)
reduce_splits_kernel(float* __restrict__ amap, float* __restrict__ bmap, float* __restrict__ gmap, int csplit, int on) {
  const int b = blockIdx.z;
  float* m = (blockIdx.y == 0 ? amap : blockIdx.y == 1 ? bmap : gmap) + (size_t)b * csplit * on;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < on; i += gridDim.x * 256) {
    float s = 0.f;
    for (int q = 0; q < csplit; ++q) s += m[(size_t)q * on + i];
    m[i] = s;
  }
}

// R1: mean / unbiased std / min / max of gamma
__global__ void __launch_bounds__(256)
gamma_stats_kernel(MapArgs a) {
  __shared__ double redd[33];
  __shared__ float redf[33];
  const int b = blockIdx.y;
  double s = 0, ss = 0;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < a.on; i += gridDim.x * 256) {
    const float gmm = map_sum(a.gmap, a.stride, a.on, b, i) / (float)a.c;
    s += gmm; ss += (double)gmm * gmm;
    mn = fminf(mn, gmm); mx = fmaxf(mx, gmm);
  }
  const double ts = block_sum<double>(s, redd), tss = block_sum<double>(ss, redd);
  // block min / max
  mn = warp_min(mn); mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { redf[threadIdx.x >> 5] = mn; }
  __syncthreads();
  float bmn = INFINITY;
  for (int q = 0; q < 8; ++q) bmn = fminf(bmn, redf[q]);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { redf[threadIdx.x >> 5] = mx; }
  __syncthreads();
  float bmx = -INFINITY;
  for (int q = 0; q < 8; ++q) bmx = fmaxf(bmx, redf[q]);
  const double vals[4] = {ts, tss, (double)bmn, (double)bmx};
  if (publish_partials(a.red, a.cnt, b, vals) && threadIdx.x == 0) {
    double S = 0, SS = 0, MN = INFINITY, MX = -INFINITY;
    for (unsigned int q = 0; q < gridDim.x; ++q) {
      const double* r = a.red + ((size_t)b * kRedBlocks + q) * 4;
      S += r[0]; SS += r[1]; MN = fmin(MN, r[2]); MX = fmax(MX, r[3]);
    }
    const double n = (double)a.on, mean = S / n;
    const double var = (SS - n * mean * mean) / (n - 1.0);   // unbiased (torch.std default); n == 1 -> NaN like torch
    LvlScal* sc = a.scal + (size_t)b * kNumLevels + a.level;
    sc->mean = (float)mean;
    sc->stdv = var != var ? NAN : (float)sqrt(fmax(var, 0.0));
    sc->gmin = (float)MN;
    sc->gmax = (float)MX;
  }
}

// per-(pair, level) constants of the min-max normalised sigmoid, formed once per thread
struct PsNorm {
  float mean, den, smin, range;
  __device__ explicit PsNorm(const LvlScal& sc) {
    mean = sc.mean;
    den = sc.stdv + 1e-12f;
    smin = sigmoidf_ref((sc.gmin - sc.mean) / den);
    range = sigmoidf_ref((sc.gmax - sc.mean) / den) - smin + 1e-12f;
  }
};
__device__ inline float ps_norm(const MapArgs& a, const PsNorm& n, int b, int i) {
  const float gmm = map_sum(a.gmap, a.stride, a.on, b, i) / (float)a.c;
  const float ps = sigmoidf_ref((gmm - n.mean) / n.den);
  return (ps - n.smin) / n.range;
}

// R2: min / max of ps_n * interp(prev)
__global__ void __launch_bounds__(256)
prod_minmax_kernel(MapArgs a) {
  __shared__ float redf[33];
  const int b = blockIdx.y;
  const PsNorm sc(a.scal[(size_t)b * kNumLevels + a.level]);
  const float isy = interp_scale(a.ph, a.oh), isx = interp_scale(a.pw, a.ow);
  float mn = INFINITY, mx = -INFINITY;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < a.on; i += gridDim.x * 256) {
    const int oy = i / a.ow, ox = i - oy * a.ow;
    const float pr = ps_norm(a, sc, b, i) * interp_prev(a.prev, a.ph, a.pw, b, oy, ox, isy, isx);
    mn = fminf(mn, pr); mx = fmaxf(mx, pr);
  }
  mn = warp_min(mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mn;
  __syncthreads();
  float bmn = INFINITY;
  for (int q = 0; q < 8; ++q) bmn = fminf(bmn, redf[q]);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = mx;
  __syncthreads();
  float bmx = -INFINITY;
  for (int q = 0; q < 8; ++q) bmx = fmaxf(bmx, redf[q]);
  const double vals[4] = {(double)bmn, (double)bmx, 0.0, 0.0};
  if (publish_partials(a.red, a.cnt, b, vals) && threadIdx.x == 0) {
    double MN = INFINITY, MX = -INFINITY;
    for (unsigned int q = 0; q < gridDim.x; ++q) {
      const double* r = a.red + ((size_t)b * kRedBlocks + q) * 4;
      MN = fmin(MN, r[0]); MX = fmax(MX, r[1]);
    }
    LvlScal* sc2 = a.scal + (size_t)b * kNumLevels + a.level;
    sc2->pmin = (float)MN;
    sc2->pmax = (float)MX;
  }
}

// R3: store the normalised ps_prod map and reduce D_k = mean((1-ps) A + ps B)
__global__ void __launch_bounds__(256)
prod_store_kernel(MapArgs a) {
  __shared__ double redd[33];
  const int b = blockIdx.y;
  const LvlScal raw = a.scal[(size_t)b * kNumLevels + a.level];
  const PsNorm sc(raw);
  const float isy = interp_scale(a.ph, a.oh), isx = interp_scale(a.pw, a.ow);
  const float pmin = raw.pmin, prange = raw.pmax - raw.pmin + 1e-12f;
  double s = 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < a.on; i += gridDim.x * 256) {
    const int oy = i / a.ow, ox = i - oy * a.ow;
    const float pr = ps_norm(a, sc, b, i) * interp_prev(a.prev, a.ph, a.pw, b, oy, ox, isy, isx);
    const float p = (pr - pmin) / prange;
    a.ps_out[(size_t)b * a.on + i] = p;
    const float av = map_sum(a.amap, a.stride, a.on, b, i), bv = map_sum(a.bmap, a.stride, a.on, b, i);
    const float dv = (1.f - p) * av + p * bv;
    if (a.store_d) a.gmap[(size_t)b * a.stride * a.on + i] = dv;   // gamma of pixel i is no longer needed (same thread read it)
    s += (double)dv;
  }
  const double ts = block_sum<double>(s, redd);
  const double vals[4] = {ts, 0.0, 0.0, 0.0};
  if (publish_partials(a.red, a.cnt, b, vals) && threadIdx.x == 0) {
    double S = 0;
    for (unsigned int q = 0; q < gridDim.x; ++q) S += a.red[((size_t)b * kRedBlocks + q) * 4];
    a.scal[(size_t)b * kNumLevels + a.level].d = (float)(S / (double)a.on);
  }
}

// fallback level: ps = sigmoid(gamma), ps_prod = ps * interp(prev -> 1x1), no min-max (ADISTS.py:92-97)
__global__ void fallback_ps_kernel(MapArgs a, int pairs) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= pairs) return;
  const float gmm = a.gmap[b] / (float)a.c;
  const float ps = sigmoidf_ref(gmm);
  const float pr = ps * interp_prev(a.prev, a.ph, a.pw, b, 0, 0, 0.f, 0.f);
  a.ps_out[b] = pr;
  a.scal[(size_t)b * kNumLevels + a.level].d = (1.f - pr) * a.amap[b] + pr * a.bmap[b];
}

// (6) score = 1 - sum_k D_k, accumulated k = 5..0 like the reference loop (ADISTS.py:165-197)
__global__ void final_kernel(const LvlScal* __restrict__ scal, int pairs, float* __restrict__ score) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= pairs) return;
  float d = 0.f;
  for (int k = kNumLevels - 1; k >= 0; --k) d += scal[(size_t)b * kNumLevels + k].d;
  score[b] = 1.f - d;
}

// (7) as_map (ADISTS.py:163,188-190,192-193): 1 - sum_k bilinear(D_map_k -> (H, W), align_corners=False)
struct DMapLevels {
  const float* dmap[kNumLevels];   // windowed levels: [pairs][csplit][oh*ow], slot 0 holds D; fallback levels: unused
  int oh[kNumLevels], ow[kNumLevels], csplit[kNumLevels], windowed[kNumLevels];
};

__device__ inline float bilinear_ac_false(const float* p, int ih, int iw, int oy, int ox, int oh, int ow) {
  // torch upsample_bilinear2d, align_corners=False: src = (dst + 0.5) * in/out - 0.5, clamped at 0
  const float sy = (float)ih / (float)oh, sx = (float)iw / (float)ow;
  const float fy = fmaxf(((float)oy + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf(((float)ox + 0.5f) * sx - 0.5f, 0.f);
  const int y0 = min((int)fy, ih - 1), x0 = min((int)fx, iw - 1);
  const int y1 = min(y0 + 1, ih - 1), x1 = min(x0 + 1, iw - 1);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  return hy * (hx * p[y0 * iw + x0] + lx * p[y0 * iw + x1]) + ly * (hx * p[y1 * iw + x0] + lx * p[y1 * iw + x1]);
}

__global__ void __launch_bounds__(256)
as_map_kernel(DMapLevels lv, const LvlScal* __restrict__ scal, int H, int W, float* __restrict__ out) {
  const int b = blockIdx.y;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < H * W; i += gridDim.x * 256) {
    const int oy = i / W, ox = i - oy * W;
    float acc = 0.f;
    for (int k = kNumLevels - 1; k >= 0; --k) {
      if (lv.windowed[k])
        acc += bilinear_ac_false(lv.dmap[k] + (size_t)b * lv.csplit[k] * lv.oh[k] * lv.ow[k], lv.oh[k], lv.ow[k], oy, ox, H, W);
      else
        acc += scal[(size_t)b * kNumLevels + k].d;   // 1x1 map
    }
    out[(size_t)b * H * W + i] = 1.f - acc;
  }
}

Taps make_taps(int window) {
  // ADISTS.gaussian (ADISTS.py:102-104): exp in double, stored as fp32, normalised in fp32
  Taps t;
  const double sigma = window / 3.0;
  float sum = 0.f;
  for (int i = 0; i < kAdistsMaxWindow; ++i) t.g[i] = 0.f;
  for (int i = 0; i < window; ++i) {
    const double d = (double)(i - window / 2);
    t.g[i] = (float)exp(-(d * d) / (2.0 * sigma * sigma));
    sum += t.g[i];
  }
  for (int i = 0; i < window; ++i) t.g[i] = t.g[i] / sum;
  for (int i = 0; i < kAdistsMaxWindow; ++i) t.g2[i] = make_float2(t.g[i], t.g[i]);
  return t;
}

template <int WIN, typename TI>
int launch_window(const LevelPlan& L, const TI* fx, const TI* fy, int pairs, const double* chan, const float* wt,
                  const Taps& taps, float* amap, float* bmap, float* gmap, cudaStream_t st) {
  constexpr int IN = WT + WIN - 1;
  constexpr int LDI = IN | 1;
  constexpr size_t smem = (size_t)(2 * IN * LDI + 5 * IN * (WT + 1) + 2) * sizeof(float);
  ensure_dynamic_smem(window_kernel<WIN, TI>, (int)smem);
  dim3 grid(L.tiles_x * L.tiles_y, L.csplit, pairs);
  window_kernel<WIN, TI><<<grid, 256, smem, st>>>(fx, fy, L.c, L.h, L.w, L.oh, L.ow, L.tiles_x, L.csplit, L.choff, chan, wt,
                                             taps, amap, bmap, gmap);
  return 1;
}

}  // namespace

bool adists_window_supported(int window) { return window == 21 || window == 11 || window == 7; }

size_t adists_workspace_bytes(int pairs, int H, int W, int window) { return make_plan(pairs, H, W, window).total; }

namespace {
template <typename TI>
int level_front(const LevelPlan& L, const TI* x, const TI* y, int pairs, float* sums, cudaStream_t st) {
  dim3 grid(L.chunks, L.c, pairs);
  sums7_kernel<TI><<<grid, 256, 0, st>>>(x, y, L.c, L.hw, L.chunks, sums);
  return 1;
}
template <typename TI>
int level_entropy(const LevelPlan& L, const TI* x, int pairs, const double* chan, float* ent, cudaStream_t st) {
  dim3 grid(L.chunks, L.c, pairs);
  entropy_kernel<TI><<<grid, 256, 0, st>>>(x, L.c, L.hw, L.chunks, L.choff, chan, ent);
  return 1;
}
template <typename TI>
int level_window(int window, const LevelPlan& L, const TI* x, const TI* y, int pairs, const double* chan, const float* wt,
                 const Taps& taps, float* amap, float* bmap, float* gmap, cudaStream_t st) {
  switch (window) {
    case 21: return launch_window<21, TI>(L, x, y, pairs, chan, wt, taps, amap, bmap, gmap, st);
    case 11: return launch_window<11, TI>(L, x, y, pairs, chan, wt, taps, amap, bmap, gmap, st);
    default: return launch_window<7, TI>(L, x, y, pairs, chan, wt, taps, amap, bmap, gmap, st);
  }
}
}  // namespace

bool adists_planes_sums_supported(int level, int H, int W) {
  int h = H, w = W;
  for (int k = 2; k <= level; ++k) { h = pooled(h); w = pooled(w); }
  return level >= 1 && level_channels(level) % 64 == 0 && h * w >= kPlaneTile;
}

int adists_pool_strips_and_sums(const bf16* planes, bf16* pooled_out, int pairs, int level, int H, int W, int window, void* ws,
                                cudaStream_t st, float* dists_partials) {
  const Plan p = make_plan(pairs, H, W, window);
  const LevelPlan& L = p.lv[level];
  if (L.chunks_s <= 0) return 0;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  return launch_pool_strips_sums(planes, pooled_out, pairs, L.c, L.h, L.w, reinterpret_cast<float*>(base + L.off_sum_s),
                                 dists_partials, st);
}

int adists_planes_and_sums(const bf16* nhwc, bf16* planes, int pairs, int level, int H, int W, int window, void* ws,
                           cudaStream_t st, bool strips) {
  const Plan p = make_plan(pairs, H, W, window);
  const LevelPlan& L = p.lv[level];
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  constexpr int smem = 2 * kPlaneTile * 33 * (int)sizeof(uint32_t);
  ensure_dynamic_smem(planes_sums_kernel, smem);
  dim3 grid(L.chunks8, L.c / 64, pairs);
  planes_sums_kernel<<<grid, 256, smem, st>>>(nhwc, planes, pairs, L.c, L.hw, L.chunks8,
                                             reinterpret_cast<float*>(base + L.off_sum8), strips ? L.w : 0, L.h);
  return 1;
}

size_t adists_ps_elems(int H, int W, int window, int level) { return (size_t)make_plan(1, H, W, window).lv[level].on; }

int adists_stage(const void* const fx[6], const void* const fy[6], bool feats_bf16, int pairs, int H, int W, int window,
                 void* ws, float* score, cudaStream_t st, float* map_out, const AdistsParts* parts, const AdistsProf* prof,
                 unsigned sums_done_mask, const AdistsOptions* opt, unsigned strips_mask) {
  const AdistsOptions options = opt ? *opt : AdistsOptions();
  const bool use_mma = options.use_mma;
  int launches = 0;
  struct Part {
    const AdistsProf* p;
    const int* n;
    int n0;
    Part(const AdistsProf* pr, int part, const int* counter) : p(pr), n(counter), n0(*counter) { if (p && p->begin) p->begin(p->user, part); }
    ~Part() { if (p && p->end) p->end(p->user, *n - n0); }
  };
  // level 0 (the raw frames) is always fp32; levels 1..5 are bf16 planes when feats_bf16
  auto is16 = [&](int k) { return feats_bf16 && k > 0; };
  auto F32 = [](const void* p) { return static_cast<const float*>(p); };
  auto B16 = [](const void* p) { return static_cast<const bf16*>(p); };
  const Plan p = make_plan(pairs, H, W, window);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  double* chan = reinterpret_cast<double*>(base + p.off_chan);
  float* wt = F(p.off_wt);
  LvlScal* scal = reinterpret_cast<LvlScal*>(base + p.off_scal);
  double* red = reinterpret_cast<double*>(base + p.off_red);
  unsigned int* cnt = reinterpret_cast<unsigned int*>(base + p.off_cnt);
  const Taps taps = make_taps(window);

  cudaMemsetAsync(cnt, 0, (size_t)pairs * sizeof(unsigned int), st);

  SumLevels sl;
  EntLevels el;
  {
  Part part(prof, 0, &launches);
  for (int k = 0; k < kNumLevels; ++k) {
    const LevelPlan& L = p.lv[k];
    // sums already in the workspace: from planes_sums_kernel (bit 0..5 of the mask) or, bit 8 + k, from the strip-plane
    // pooling kernel (adists_strips.cu); same partial layout, different chunking
    const bool from_strips = ((sums_done_mask >> (8 + k)) & 1) != 0 && L.chunks_s > 0;
    const bool fused = from_strips || (((sums_done_mask >> k) & 1) != 0 && L.chunks8 > 0);
    sl.hw[k] = L.hw;
    sl.fused[k] = fused ? 1 : 0;
    if (fused) {
      sl.partials[k] = from_strips ? F(L.off_sum_s) : F(L.off_sum8);
      sl.chunks[k] = from_strips ? L.chunks_s : L.chunks8;
      el.partials[k] = nullptr; el.chunks[k] = 0;
      continue;
    }
    launches += is16(k) ? level_front<bf16>(L, B16(fx[k]), B16(fy[k]), pairs, F(L.off_sum), st)
                        : level_front<float>(L, F32(fx[k]), F32(fy[k]), pairs, F(L.off_sum), st);
    sl.partials[k] = F(L.off_sum); sl.chunks[k] = L.chunks;
    el.partials[k] = F(L.off_ent); el.chunks[k] = L.chunks;
  }
  chan_kernel<<<dim3(ceil_div(kTotalChannels, 32), pairs), 256, 0, st>>>(sl, chan);
  ++launches;
  for (int k = 0; k < kNumLevels; ++k) {
    const LevelPlan& L = p.lv[k];
    if (sl.fused[k]) continue;
    launches += is16(k) ? level_entropy<bf16>(L, B16(fx[k]), pairs, chan, F(L.off_ent), st)
                        : level_entropy<float>(L, F32(fx[k]), pairs, chan, F(L.off_ent), st);
  }
  weights_kernel<<<pairs, 256, 0, st>>>(el, chan, wt);
  ++launches;
  }

  int csplit_used[kNumLevels];
  {
  Part part(prof, 1, &launches);
  for (int k = 0; k < kNumLevels; ++k) {
    const LevelPlan& L = p.lv[k];
    csplit_used[k] = L.csplit;
    int mma_launches = 0;
    const bool strips = ((strips_mask >> k) & 1) != 0;
    if (strips && !(L.windowed && is16(k) && window_mma_supported(window, L.h, L.w, fx[k], fy[k]))) return -1;
    if (L.windowed && is16(k) && (use_mma || strips) && window_mma_supported(window, L.h, L.w, fx[k], fy[k]))
      // bf16 planes: the two gaussian passes as banded-Toeplitz products on the tensor pipe (adists_tc.cu); 0 = no
      // tensor map for these planes, the CUDA-core kernel serves the level
      mma_launches = launch_window_mma(B16(fx[k]), B16(fy[k]), pairs, L.c, L.h, L.w, window, L.csplit_mma, L.choff, chan, wt,
                                       taps.g, options.mma_terms == 2 ? -1 : options.one_term_min_pixels, strips, F(L.off_a), F(L.off_b), F(L.off_g), st);
    if (mma_launches < 0 || (strips && mma_launches == 0)) return -1;
    if (mma_launches > 0) {
      csplit_used[k] = L.csplit_mma;
      launches += mma_launches;
    } else if (L.windowed) {
      launches += is16(k) ? level_window<bf16>(window, L, B16(fx[k]), B16(fy[k]), pairs, chan, wt, taps, F(L.off_a), F(L.off_b),
                                               F(L.off_g), st)
                          : level_window<float>(window, L, F32(fx[k]), F32(fy[k]), pairs, chan, wt, taps, F(L.off_a), F(L.off_b),
                                                F(L.off_g), st);
    } else {
      global_level_kernel<<<pairs, 256, 0, st>>>(L.c, L.choff, chan, wt, F(L.off_a), F(L.off_b), F(L.off_g));
      ++launches;
    }
  }
  }

  Part part(prof, 0, &launches);
  const float* prev = nullptr;
  int ph = 1, pw = 1;
  for (int k = kNumLevels - 1; k >= 0; --k) {
    const LevelPlan& L = p.lv[k];
    MapArgs a;
    a.amap = F(L.off_a); a.bmap = F(L.off_b); a.gmap = F(L.off_g);
    if (L.windowed && csplit_used[k] > 1) {
      int nb = ceil_div(L.on, 256);
      if (nb > 1024) nb = 1024;
      reduce_splits_kernel<<<dim3(nb, 3, pairs), 256, 0, st>>>(F(L.off_a), F(L.off_b), F(L.off_g), csplit_used[k], L.on);
      ++launches;
    }
    a.csplit = 1; a.stride = csplit_used[k]; a.on = L.on; a.oh = L.oh; a.ow = L.ow; a.c = L.c;
    a.prev = prev; a.ph = ph; a.pw = pw;
    a.ps_out = F(L.off_ps); a.scal = scal; a.level = k; a.red = red; a.cnt = cnt;
    a.store_d = map_out ? 1 : 0;
    if (L.windowed) {
      int nblk = ceil_div(L.on, 256 * 2);
      if (nblk > kRedBlocks) nblk = kRedBlocks;
      if (nblk < 1) nblk = 1;
      dim3 grid(nblk, pairs);
      gamma_stats_kernel<<<grid, 256, 0, st>>>(a);
      prod_minmax_kernel<<<grid, 256, 0, st>>>(a);
      prod_store_kernel<<<grid, 256, 0, st>>>(a);
      launches += 3;
    } else {
      fallback_ps_kernel<<<ceil_div(pairs, 128), 128, 0, st>>>(a, pairs);
      ++launches;
    }
    prev = a.ps_out; ph = L.oh; pw = L.ow;
  }
  final_kernel<<<ceil_div(pairs, 128), 128, 0, st>>>(scal, pairs, score);
  ++launches;
  if (parts) {
    if (parts->weights)
      cudaMemcpyAsync(parts->weights, wt, (size_t)pairs * kTotalChannels * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (parts->level_terms)
      cudaMemcpy2DAsync(parts->level_terms, sizeof(float), &scal[0].d, sizeof(LvlScal), sizeof(float),
                        (size_t)pairs * kNumLevels, cudaMemcpyDeviceToDevice, st);
    for (int k = 0; k < kNumLevels; ++k)
      if (parts->ps[k])
        cudaMemcpyAsync(parts->ps[k], F(p.lv[k].off_ps), (size_t)pairs * p.lv[k].on * sizeof(float), cudaMemcpyDeviceToDevice, st);
  }
  if (map_out) {
    DMapLevels dl;
    for (int k = 0; k < kNumLevels; ++k) {
      const LevelPlan& L = p.lv[k];
      dl.dmap[k] = F(L.off_g); dl.oh[k] = L.oh; dl.ow[k] = L.ow; dl.csplit[k] = csplit_used[k]; dl.windowed[k] = L.windowed ? 1 : 0;
    }
    int nblk = ceil_div(H * W, 256 * 4);
    if (nblk > 2048) nblk = 2048;
    dim3 grid(nblk, pairs);
    as_map_kernel<<<grid, 256, 0, st>>>(dl, scal, H, W, map_out);
    ++launches;
  }
  return launches;
}

}  // namespace nqa
