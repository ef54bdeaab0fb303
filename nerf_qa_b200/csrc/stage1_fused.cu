// VGG stage 1 in ONE kernel: input normalisation -> conv1_1 + bias + ReLU -> conv1_2 + bias + ReLU.
//
// relu1_1 (64 x H x W per image, the largest activation of the trunk) never leaves the SM: per 16 x 8 output
// tile the kernel
//   1. producer warps stage the normalised 20 x 12 x 3 input patch in shared memory and build the im2col rows
//      (K = 27 padded to 32, bf16, canonical no-swizzle K-major layout) of the 18 x 10 halo of conv1_1 outputs
//      that conv1_2 needs;
//   2. one thread issues 4 tcgen05.mma (2 x M=128, N=64, K=2x16) -> 256 x 64 fp32 in TMEM (180 rows used);
//   3. six "mid" epilogue warps (four for halo rows 0..127, two for rows 128..179, so that both accumulator halves
//      drain in parallel) read that accumulator, add bias, ReLU, zero the pixels outside the image (the zero padding
//      of conv1_2), round to bf16 and write the halo tile into shared memory as a K-major UMMA operand: linear
//      [8-channel chunk][halo row][16 B] by default, or the 128-byte-swizzled layout TMA would have produced;
//   4. the same MMA thread issues the 36 tcgen05.mma of conv1_2 (9 taps x 4 k-steps, M=128, N=64): the A operand of
//      tap (r, s) is the halo tile viewed through a descriptor shifted by (10 r + s) rows, 8-row groups 10 rows apart;
//      the 72 KB of conv1_2 weights stay resident in shared memory (TMA, loaded once per CTA);
//   5. four "final" epilogue warps: tcgen05.ld -> +bias -> ReLU -> bf16 -> swizzled staging tile -> one TMA store.
// Everything is double-buffered (im2col rows, both accumulators, halo tile) and handed over with mbarriers, so the
// conv1_1 MMAs + mid epilogue of tile i+1 overlap the conv1_2 MMAs of tile i.  Persistent grid = #SMs.
//
// PAIR = true runs the same pipeline on a cluster of two CTAs (one TPC): each CTA keeps its own tile, producers and
// epilogues, but every tcgen05.mma is a cta_group::2 instruction issued by the leader for both tiles (M = 256) and each
// CTA holds only half of the conv1_1 / conv1_2 weights (32 of the 64 output channels): 128 + 32 operand rows fetched
// per SM and MMA instead of 128 + 64 (2.46 vs 2.72 ms per 16 frames, 54% vs 48% tensor-pipe activity; DESIGN.md §4 has
// the history -- the pair only paid off once the producers and the mid epilogue were no longer the limiters).
// Barriers the MMA warp waits on live in the leader CTA (the follower arrives remotely), the ones it signals are
// multicast to both CTAs.
//
// Reference semantics: DISTS_pt.py:92 (normalise), vgg16.features[0:4] (DISTS_pt.py:36-37), ADISTS.py:113-114.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace nqa {
namespace s1 {

using namespace tc;

constexpr int TH = 16, TW = 8;                 // output tile (M = 128, m = 8*row + col)
constexpr int HH = TH + 2, HW = TW + 2;        // conv1_1 halo tile 18 x 10
constexpr int HROWS = HH * HW;                 // 180 im2col rows
constexpr int IH = TH + 4, IW = TW + 4;        // input patch 20 x 12
constexpr int IN_ELEMS = 3 * IH * IW;          // 720
constexpr int KPAD = 32, C = 64;
constexpr int A1_BYTES = 256 * KPAD * 2;       // 16 KB (two M=128 operands)
constexpr int A2_BYTES = 23 * 1024;            // 180 rows x 128 B, rounded to the 1024-byte swizzle atom
constexpr int W2_TILE = C * 64 * 2;            // 8 KB per tap
constexpr int W1_BYTES = C * KPAD * 2;         // 4 KB
constexpr int NUM_PRODUCERS = 192;             // warps 0..5
constexpr int MMA_WARP = 6;                    // warp 6: TMEM alloc, weight TMA, MMA issue
// mid epilogue: warps 7..10 take halo rows 0..127 (TMEM lane quarters 3,0,1,2), warps 12 and 13 rows 128..179
// (quarters 0 and 1 of the second accumulator half); warp 11 would own an empty quarter and idles.  Final epilogue =
// warps 14..17 (distinct warp%4 each: a warp may only read the TMEM lanes 32*(warp%4)..+31).
constexpr int FIN_WARP0 = 14;
constexpr int NUM_MID_WARPS = 6;
constexpr int NUM_THREADS = 18 * 32;
constexpr uint32_t TMEM_COLS = 512;            // D1: 2 x 128 columns, D2: 2 x 64 columns at 256

constexpr int OFF_A2 = 0;
constexpr int OFF_W2 = OFF_A2 + 2 * A2_BYTES;
constexpr int OFF_A1 = OFF_W2 + 9 * W2_TILE;
constexpr int OFF_W1 = OFF_A1 + 2 * A1_BYTES;
constexpr int OFF_IN = OFF_W1 + W1_BYTES;
constexpr int OFF_STG = (OFF_IN + 2 * IN_ELEMS * 4 + 1023) & ~1023;   // 2 x 16 KB output staging for the TMA store
constexpr int OFF_BIAS = OFF_STG + 2 * 128 * 128;
constexpr int OFF_BARS = OFF_BIAS + 2 * C * 4;
constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024 /*align slack*/;

struct Params {
  const float* x;
  const float* y;
  int nx, n, h, w;
  int tiles_w, tiles_h, num_tiles;
  const bf16* w1img;   // conv1_1 weights, shared-memory image [64 x 32] no-swizzle K-major
  const float* bias1;
  const float* bias2;
  bf16* out;           // relu1_2, NHWC [n, h, w, 64]
  int strips;          // 1: relu1_2 leaves as strip planes [n][64][w/8][h][8] (kernels.h make_strip_out_map)
};

template <bool PAIR, bool A2_LINEAR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
stage1_fused_kernel(const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out,
                    const Params p) {
  // work units: single CTA -> one tile per unit; PAIR -> two consecutive tiles per unit, tile 2u + rank for this CTA
  // (1-D grid with cluster dims (2,1,1): rank = blockIdx.x & 1, derived from blockIdx so that it is warp-uniform)
  const uint32_t rank = PAIR ? (blockIdx.x & 1) : 0;
  const bool leader = rank == 0;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_units = PAIR ? (p.num_tiles + 1) / 2 : p.num_tiles;
  auto tile_of = [&](int unit) { return PAIR ? 2 * unit + (int)rank : unit; };
  constexpr int NW = PAIR ? 2 : 1;                    // CTAs whose threads arrive on the leader's barriers
  constexpr int CH = PAIR ? C / 2 : C;                // output channels whose weights this CTA holds
  constexpr int W2_SLICE = CH * 64 * 2;               // bytes per tap
  // arrive on a barrier the (leader's) MMA warp waits on
  auto arrive_mma = [&](uint64_t* bar) {
    if (PAIR) mbar_arrive_cluster(map_to_rank(bar, 0));
    else mbar_arrive(bar);
  };
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET from the __shared__ array (a round trip through uintptr_t would turn every access into a generic
  // LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* a2_s = smem + OFF_A2;
  uint8_t* w2_s = smem + OFF_W2;
  uint8_t* a1_s = smem + OFF_A1;
  uint8_t* w1_s = smem + OFF_W1;
  float* in_s = reinterpret_cast<float*>(smem + OFF_IN);
  float* bias1_s = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* bias2_s = bias1_s + C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* a1_full = bars;          // [2] 180 producer arrivals
  uint64_t* a1_empty = bars + 2;     // [2] tcgen05.commit
  uint64_t* d1_full = bars + 4;      // [2] tcgen05.commit
  uint64_t* d1_empty = bars + 6;     // [2] 4 mid-warp arrivals
  uint64_t* a2_full = bars + 8;      // [2] 128 mid-thread arrivals
  uint64_t* a2_empty = bars + 10;    // [2] tcgen05.commit
  uint64_t* d2_full = bars + 12;     // [2] tcgen05.commit
  uint64_t* d2_empty = bars + 14;    // [2] 4 final-warp arrivals
  uint64_t* w_full = bars + 16;      // conv1_2 weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // conv1_1 weights: rows [rank*CH, rank*CH + CH) of the [64 x 32] no-swizzle image (8-row groups of 512 bytes)
  for (int i = threadIdx.x; i < (CH * KPAD * 2) / 16; i += NUM_THREADS)
    reinterpret_cast<uint4*>(w1_s)[i] = reinterpret_cast<const uint4*>(p.w1img)[rank * ((CH * KPAD * 2) / 16) + i];
  if (threadIdx.x < C) {
    bias1_s[threadIdx.x] = p.bias1[threadIdx.x];
    bias2_s[threadIdx.x] = p.bias2[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_w2);
    prefetch_tmap(&tmap_out);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a1_full[s], NW * (NUM_PRODUCERS / 32));
      mbar_init(&a1_empty[s], 1);
      mbar_init(&d1_full[s], 1);
      mbar_init(&d1_empty[s], NW * NUM_MID_WARPS);
      mbar_init(&a2_full[s], NW * NUM_MID_WARPS);
      mbar_init(&a2_empty[s], 1);
      mbar_init(&d2_full[s], 1);
      mbar_init(&d2_empty[s], NW * 4);
    }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // w1 written through the generic proxy
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();              // barrier inits, w1 and the TMEM allocation of both CTAs are in place
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Tile walker: (image, tile row, tile column) of this CTA's tiles without a division per tile -- the per-tile
  // stride is decomposed once and added with carries.
  struct Walk {
    int tile, tw, th, img;
  };
  const int tile_first = tile_of(unit0), tile_step = PAIR ? 2 * unit_stride : unit_stride;
  const int step_tw = tile_step % p.tiles_w, step_th = (tile_step / p.tiles_w) % p.tiles_h,
            step_img = tile_step / (p.tiles_w * p.tiles_h);
  auto walk_begin = [&]() {
    Walk k;
    k.tile = tile_first;
    k.tw = tile_first % p.tiles_w;
    const int t = tile_first / p.tiles_w;
    k.th = t % p.tiles_h;
    k.img = t / p.tiles_h;
    return k;
  };
  auto walk_next = [&](Walk& k) {
    k.tile += tile_step;
    k.tw += step_tw;
    const int c0 = k.tw >= p.tiles_w ? 1 : 0;
    k.tw -= c0 ? p.tiles_w : 0;
    k.th += step_th + c0;
    const int c1 = k.th >= p.tiles_h ? 1 : 0;
    k.th -= c1 ? p.tiles_h : 0;
    k.img += step_img + c1;
  };

  if (warp < MMA_WARP) {
    // =============================== im2col producers ===============================
    const int pt = threadIdx.x;
    constexpr int LPT = (IN_ELEMS + NUM_PRODUCERS - 1) / NUM_PRODUCERS;   // 4 patch elements per thread
    // tile-independent part of the patch addressing: element e = (ci, rr, cc) of the 3 x 20 x 12 patch
    int e_rr[LPT], e_cc[LPT];
    size_t e_plane[LPT];
    float e_mean[LPT], e_inv[LPT];
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
      const int e = pt + j * NUM_PRODUCERS;
      const int ci = e < IN_ELEMS ? e / (IH * IW) : 0;
      const int rem = e - ci * (IH * IW);
      e_rr[j] = e < IN_ELEMS ? rem / IW : -100000;          // never inside the image
      e_cc[j] = rem - (rem / IW) * IW;
      e_plane[j] = (size_t)ci * p.h * p.w;
      e_mean[j] = ci == 0 ? 0.485f : ci == 1 ? 0.456f : 0.406f;
      e_inv[j] = ci == 0 ? kInvStd0 : ci == 1 ? kInvStd1 : kInvStd2;
    }
    // fetch() only ISSUES the global loads (raw values, no arithmetic on them), so they stay in flight while the
    // current tile's rows are built; stash() normalises and writes them to shared memory one tile later.
    float pre[LPT];
    bool pre_ok[LPT];
    auto fetch = [&](const Walk& k) {
      int img = k.img;
      const int h0 = k.th * TH, w0 = k.tw * TW;
      const bool tile_ok = k.tile < p.num_tiles;        // PAIR: the last unit of an odd tile count carries a padding tile
      if (!tile_ok) img = p.n - 1;
      const float* src = img < p.nx ? p.x + (size_t)img * 3 * p.h * p.w : p.y + (size_t)(img - p.nx) * 3 * p.h * p.w;
      // Unconditional loads from a clamped address: a load under a branch makes the compiler wait for it at the
      // reconvergence point, which serialises the four global-load latencies of a tile.
#pragma unroll
      for (int j = 0; j < LPT; ++j) {
        const int gy = h0 - 2 + e_rr[j], gx = w0 - 2 + e_cc[j];
        pre_ok[j] = tile_ok && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w;
        const int cy = min(max(gy, 0), p.h - 1), cx = min(max(gx, 0), p.w - 1);
        pre[j] = __ldg(src + e_plane[j] + (size_t)cy * p.w + cx);
      }
    };
    auto stash = [&](int buf) {
#pragma unroll
      for (int j = 0; j < LPT; ++j) {
        const int e = pt + j * NUM_PRODUCERS;
        // zero padding applies to the NORMALISED tensor
        if (e < IN_ELEMS) in_s[buf * IN_ELEMS + e] = pre_ok[j] ? (pre[j] - e_mean[j]) * e_inv[j] : 0.f;
      }
    };
    const int hr = pt / HW, hc = pt - hr * HW;
    int unit = unit0;
    Walk wk = walk_begin();
    if (unit < num_units) {
      fetch(wk);
      stash(0);
    }
    named_bar_sync(1, NUM_PRODUCERS);
    for (int local = 0; unit < num_units; ++local, unit += unit_stride) {
      const int s = local & 1;
      const uint32_t ph = (local >> 1) & 1;
      const int next = unit + unit_stride;
      if (next < num_units) {                       // global loads in flight while this tile's rows are built
        walk_next(wk);
        fetch(wk);
      }
      if (pt < HROWS) {
        const float* src = in_s + s * IN_ELEMS + hr * IW + hc;
        float v[KPAD];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) v[(ci * 3 + r) * 3 + q] = src[ci * (IH * IW) + r * IW + q];
#pragma unroll
        for (int k = 27; k < KPAD; ++k) v[k] = 0.f;
        mbar_wait(&a1_empty[s], ph ^ 1);
        uint8_t* dst = a1_s + s * A1_BYTES;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          uint4 t;
          t.x = Vec<bf16>::pack2(v[kc * 8 + 0], v[kc * 8 + 1]);
          t.y = Vec<bf16>::pack2(v[kc * 8 + 2], v[kc * 8 + 3]);
          t.z = Vec<bf16>::pack2(v[kc * 8 + 4], v[kc * 8 + 5]);
          t.w = Vec<bf16>::pack2(v[kc * 8 + 6], v[kc * 8 + 7]);
          *reinterpret_cast<uint4*>(dst + noswz_offset_bytes(pt, kc)) = t;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncwarp();
      if (lane == 0) arrive_mma(&a1_full[s]);        // one arrival per producer warp
      if (next < num_units) stash(s ^ 1);
      named_bar_sync(1, NUM_PRODUCERS);
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA issuer (whole warp, one elected lane issues; leader CTA only) ===============
    // conv1_2 weights of this CTA's channels: box {64 k, CH rows} per tap; PAIR: both CTAs' bytes are counted on the
    // leader's barrier
    if (elect_one()) {
      if (PAIR) {
        if (leader) mbar_expect_tx(w_full, 2 * 9 * W2_SLICE);
        const uint32_t bar = map_to_rank(w_full, 0);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d_pair(w2_s + tap * W2_SLICE, &tmap_w2, bar, tap * C, (int)rank * CH);
      } else {
        mbar_expect_tx(w_full, 9 * W2_SLICE);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(w2_s + tap * W2_SLICE, &tmap_w2, w_full, tap * C, 0);
      }
    }
    __syncwarp();
    if (leader) {
      constexpr uint32_t idesc = make_idesc(PAIR ? 256 : 128, C);
      const uint64_t w1_desc = make_desc_noswz(smem_u32(w1_s));
      const uint64_t w2_desc = make_desc_sw128(smem_u32(w2_s), 1024, 0);
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
        if (PAIR) umma_bf16_pair(d, a, b, idesc, acc);
        else umma_bf16(d, a, b, idesc, acc);
      };
      auto commit = [&](uint64_t* bar) {
        if (PAIR) umma_commit_pair(bar);
        else umma_commit(bar);
      };
      auto conv1_1 = [&](int local) {
        const int s = local & 1;
        const uint32_t ph = (local >> 1) & 1;
        mbar_wait(&d1_empty[s], ph ^ 1);
        mbar_wait(&a1_full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a1_desc = make_desc_noswz(smem_u32(a1_s + s * A1_BYTES));
#pragma unroll
          for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int k = 0; k < KPAD / 16; ++k)   // descriptor address field counts 16-byte units
              mma(tmem_base + s * 128 + half * 64, a1_desc + (uint64_t)((half * (A1_BYTES / 2) + k * 256) >> 4),
                  w1_desc + (uint64_t)((k * 256) >> 4), k > 0 ? 1u : 0u);
          commit(&a1_empty[s]);
          commit(&d1_full[s]);
        }
        __syncwarp();
      };
      int n_local = 0;
      for (int unit = unit0; unit < num_units; unit += unit_stride) ++n_local;
      if (n_local > 0) conv1_1(0);
      mbar_wait(w_full, 0);
      tc_fence_after();
      for (int local = 0; local < n_local; ++local) {
        if (local + 1 < n_local) conv1_1(local + 1);
        const int s = local & 1;
        const uint32_t ph = (local >> 1) & 1;
        mbar_wait(&d2_empty[s], ph ^ 1);
        mbar_wait(&a2_full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          // A2_LINEAR: relu1_1 halo tile stored without swizzle as [8-channel chunk][halo row][16 B]: a K=16 slice of
          // an 8-pixel row group is two contiguous 128-byte lines (LBO = HROWS*16 between chunks, SBO = HW*16 between
          // pixel rows of the tile); a tap shift is still a plain start-address offset.
          const uint64_t a2_desc = A2_LINEAR
              ? make_desc_noswz_strided(smem_u32(a2_s + s * A2_BYTES), HROWS * 16, HW * 16)
              : make_desc_sw128(smem_u32(a2_s + s * A2_BYTES), HW * 128, 0);
          const uint32_t d2 = tmem_base + 256 + s * C;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, q = tap - r * 3;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma(d2, a2_desc + (uint64_t)(A2_LINEAR ? (r * HW + q) + k * 2 * HROWS : (r * HW + q) * 8 + k * 2), w2_desc + (uint64_t)((tap * W2_SLICE) / 16 + k * 2),
                  (tap > 0 || k > 0) ? 1u : 0u);
          }
          commit(&a2_empty[s]);
          commit(&d2_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp < FIN_WARP0) {
    // =============================== mid epilogue: relu1_1 halo tile -> shared memory ===============================
    const int quarter = warp & 3;
    const int half = warp >= 11 ? 1 : 0;
    int local = 0;
    Walk wk = walk_begin();
    if (warp != 11)
    for (int unit = unit0; unit < num_units; unit += unit_stride, ++local, walk_next(wk)) {
      const int s = local & 1;
      const uint32_t ph = (local >> 1) & 1;
      const int h0 = wk.th * TH, w0 = wk.tw * TW;
      mbar_wait(&d1_full[s], ph);
      mbar_wait(&a2_empty[s], ph ^ 1);
      tc_fence_after();
      uint8_t* a2 = a2_s + s * A2_BYTES;
      {
        const int row = half * 128 + quarter * 32 + lane;
        const int rr = row / HW, cc = row - rr * HW;
        const int gy = h0 - 1 + rr, gx = w0 - 1 + cc;
        const bool inside = row < HROWS && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + s * 128 + half * 64 + ch * 32, r);
          tmem_ld_wait();
          if (row < HROWS) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                v[e] = inside ? fmaxf(__uint_as_float(r[j * 8 + e]) + bias1_s[ch * 32 + j * 8 + e], 0.f) : 0.f;
              uint4 t;
              t.x = Vec<bf16>::pack2(v[0], v[1]);
              t.y = Vec<bf16>::pack2(v[2], v[3]);
              t.z = Vec<bf16>::pack2(v[4], v[5]);
              t.w = Vec<bf16>::pack2(v[6], v[7]);
              const int chunk = ch * 4 + j;                      // 16-byte chunk = 8 channels
              *reinterpret_cast<uint4*>(a2 + (A2_LINEAR ? chunk * (HROWS * 16) + row * 16
                                                        : row * 128 + ((chunk ^ (row & 7)) << 4))) = t;
            }
          }
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
      __syncwarp();
      if (lane == 0) {
        arrive_mma(&d1_empty[s]);
        arrive_mma(&a2_full[s]);
      }
    }
  } else {
    // =============================== final epilogue: relu1_2 -> shared memory -> TMA store ===============================
    // Each thread owns one pixel (TMEM lane) and its 64 channels = one 128-byte row; rows are staged in the
    // 128-byte-swizzled layout of the output tensor map and written by ONE TMA store per tile (full 128-byte lines,
    // image-edge clipping by the TMA unit) instead of 32 scattered 16-byte stores per warp instruction.
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const bool elected = warp == FIN_WARP0 && lane == 0;
    uint8_t* stg_s = smem + OFF_STG;
    int local = 0;
    Walk wk = walk_begin();
    for (int unit = unit0; unit < num_units; unit += unit_stride, ++local, walk_next(wk)) {
      const int s = local & 1;
      const uint32_t ph = (local >> 1) & 1;
      const int img = wk.img, h0 = wk.th * TH, w0 = wk.tw * TW;   // a padding tile has img == n: the TMA store clips it away
      mbar_wait(&d2_full[s], ph);
      tc_fence_after();
      if (elected) tma_store_wait_read<1>();          // the store issued two tiles ago has read staging buffer s
      named_bar_sync(2, 128);
      uint8_t* row = stg_s + s * (128 * 128) + m * 128;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + 256 + s * C + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaxf(__uint_as_float(r[j * 8 + e]) + bias2_s[ch * 32 + j * 8 + e], 0.f);
          uint4 t;
          t.x = Vec<bf16>::pack2(v[0], v[1]);
          t.y = Vec<bf16>::pack2(v[2], v[3]);
          t.z = Vec<bf16>::pack2(v[4], v[5]);
          t.w = Vec<bf16>::pack2(v[6], v[7]);
          if (p.strips) {
            // [channel][pixel]: a warp's 32 pixels of one channel are 64 contiguous bytes (conflict-free 2-byte stores)
            unsigned short* col = reinterpret_cast<unsigned short*>(stg_s + s * (128 * 128)) + (ch * 32 + j * 8) * 128 + m;
            const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              col[(2 * e) * 128] = (unsigned short)(tw[e] & 0xffffu);
              col[(2 * e + 1) * 128] = (unsigned short)(tw[e] >> 16);
            }
          } else {
            *reinterpret_cast<uint4*>(row + (((ch * 4 + j) ^ (m & 7)) << 4)) = t;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_mma(&d2_empty[s]);
      fence_proxy_async_smem();
      named_bar_sync(3, 128);
      if (elected) {
        if (p.strips) tma_store_4d(&tmap_out, stg_s + s * (128 * 128), h0 * 8, w0 >> 3, 0, img);
        else tma_store_4d(&tmap_out, stg_s + s * (128 * 128), 0, w0, h0, img);
        tma_store_commit();
      }
    }
    if (elected) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();              // the peer may still signal barriers in this CTA's shared memory
  if (warp == MMA_WARP) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace s1

// w2map: 2-D tensor map over the conv1_2 weights [64][9*64] bf16, SWIZZLE_128B, with box {64, 64} (single CTA) or
// {64, 32} (pair = true: cluster of two CTAs, each holding half of the output channels).  Returns the number of launches
// (0 if the output tensor map cannot be encoded or the launch fails).
int launch_stage1_fused(const float* x, const float* y, int nx, int n, int h, int w, const bf16* w1img,
                        const float* bias1, const CUtensorMap& w2map, const float* bias2, bf16* out, int sm_count, bool pair,
                        bool a2_linear, cudaStream_t st, bool strips_out) {
  CUtensorMap outmap;
  if (strips_out ? make_strip_out_map(&outmap, out, n, s1::C, h, w) != 0
                 : conv_tc_make_act_map(&outmap, out, n, h, w, s1::C, s1::TW, s1::TH) != 0)
    return 0;
  using namespace s1;
  Params p;
  p.x = x; p.y = y; p.nx = nx; p.n = n; p.h = h; p.w = w;
  p.tiles_w = ceil_div(w, TW);
  p.tiles_h = ceil_div(h, TH);
  p.num_tiles = n * p.tiles_h * p.tiles_w;
  p.w1img = w1img; p.bias1 = bias1; p.bias2 = bias2; p.out = out;
  p.strips = strips_out ? 1 : 0;
  if (pair) {
    auto kern = a2_linear ? stage1_fused_kernel<true, true> : stage1_fused_kernel<true, false>;
    ensure_dynamic_smem(kern, SMEM_BYTES);
    const int units = (p.num_tiles + 1) / 2;
    int clusters = sm_count / 2;
    if (units < clusters) clusters = units;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, w2map, outmap, p) == cudaSuccess ? 1 : 0;
  }
  auto kern = a2_linear ? stage1_fused_kernel<false, true> : stage1_fused_kernel<false, false>;
  ensure_dynamic_smem(kern, SMEM_BYTES);
  const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
  kern<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(w2map, outmap, p);
  return 1;
}

}  // namespace nqa
