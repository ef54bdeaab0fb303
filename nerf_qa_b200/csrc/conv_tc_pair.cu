// CTA-pair (tcgen05 cta_group::2) version of the halo-reuse 3x3 convolution.
//
// Measured on B200 (profiles/r1h_*): a single-CTA SS-mode tcgen05.mma fetches its shared-memory operands at
// ~64 B/clk, i.e. an M=128 x N x K=16 instruction costs (128 + N)/2 cycles of operand traffic against
// 128*N/256 cycles of math -- N=256 tiles top out at ~65% of the tensor pipe, N=128 at 50%.  Pairing two CTAs of
// one TPC (cluster 2x1x1) lets ONE tcgen05.mma.cta_group::2 compute a 256 x N tile: each CTA supplies its own
// 128-pixel A tile but only HALF of the weight tile (N/2 rows), so the per-SM operand traffic drops to
// (128 + N/2)/2 cycles per instruction -- balanced with the math at N=256.
//
//   * tile pair = two consecutive 16 x 8 pixel tiles (CTA rank r takes m-tile 2t + r), same channel tile;
//   * A: per 64-channel chunk one 4-D TMA box {64, HBW, 18, 1} (halo reuse across the 9 taps, as in
//     conv_tc_halo.cu); B: per tap each CTA loads rows [n0 + r*N/2, n0 + (r+1)*N/2) of the weight tile, either
//     once for the whole kernel (resident) or through an mbarrier ring;
//   * every TMA load (both CTAs) signals the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2 with
//     the mapa-translated barrier address); the leader's elected lane issues all MMAs and releases smem slots /
//     publishes accumulators to BOTH CTAs with tcgen05.commit ... .multicast::cluster;
//   * accumulators form a ring over all 512 TMEM columns (2 x N=256 or 4 x N=128), so the epilogue latency of the short
//     N=128 tiles (cross-CTA hand-offs included) stays hidden behind the MMAs of the next tiles;
//   * each CTA's epilogue warps read their own 128 TMEM lanes, add bias, ReLU, stage through swizzled shared
//     memory and TMA-store their own tile; the follower's warps hand the accumulator back with a remote
//     mbarrier.arrive on the leader's tmem_empty barrier.
//
//
// SPLIT (bf16x3 trunk, forward only): activations and weights arrive as bf16 hi + lo pairs; a stage holds the hi and the
// lo halo tile, a weight slot the hi and the lo half-tile, every k-step issues hi*hi + hi*lo + lo*hi into the same fp32
// accumulator, and the epilogue writes the result as hi = bf16(v), lo = bf16(v - hi) through two staging tiles and two
// TMA stores.  Same halo reuse and CTA pairing as the bf16 path: the per-tap kernel of conv_tc.cu re-fetches the A tile
// nine times per chunk and is bound by L2 bandwidth at three MMAs per k-step.
//
// Reference semantics: torchvision vgg16.features convs (DISTS_pt.py:36-49, ADISTS.py:44-57).
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace nqa {
namespace tcp {

using namespace tc;

constexpr int BLOCK_K = 64, UMMA_K = 16;
constexpr int TH = 16, TW = 8, HALO_H = TH + 2;
constexpr int NUM_THREADS = 192;
constexpr int STG_BYTES = 2 * 128 * 128;

struct Params {
  int n, h, w, cin, cout;
  int tiles_w, tiles_h, num_n_tiles, num_m_pairs, num_items;
  int hbw, a_stage_bytes;
  const float* bias;
  int epi;             // ConvEpilogue
  const bf16* mask;    // kEpiMask: saved activation of the layer below, [n,h,w,cout]
  int strips;          // 1: the output goes out as strip planes (kernels.h make_strip_out_map), kEpiBiasRelu only
};

template <int BLOCK_N, bool RESIDENT, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv3x3_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_a_lo,
                       const __grid_constant__ CUtensorMap tmap_b_lo, const __grid_constant__ CUtensorMap tmap_out_lo, const Params p,
                       int a_stages, int b_stages) {
  static_assert(!(SPLIT && RESIDENT), "hi + lo weights are streamed");
  constexpr int HALF_N = BLOCK_N / 2;
  constexpr int NSPLIT = SPLIT ? 2 : 1;                            // hi (+ lo) tiles per stage / slot
  constexpr int ACC = 512 / BLOCK_N;                              // accumulator ring in TMEM: 2 x 256 or 4 x 128 columns
  constexpr int ACC_LOG2 = ACC == 4 ? 2 : 1;
  constexpr int B_TILE = HALF_N * BLOCK_K * 2;                   // this CTA's half of one tap's weight tile
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET from the __shared__ array (a round trip through uintptr_t would turn every access into a generic
  // LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int chunks = p.cin / BLOCK_K;
  uint8_t* stg_smem = smem;
  uint8_t* a_smem = smem + STG_BYTES;
  const int a_stride = NSPLIT * p.a_stage_bytes;                  // bytes per A stage (hi halo tile, then lo)
  constexpr int B_SLOT = NSPLIT * B_TILE;                          // bytes per weight slot (hi half-tile, then lo)
  uint8_t* b_smem = a_smem + a_stages * a_stride;
  const int b_slots = RESIDENT ? chunks * 9 : b_stages;
  uint8_t* tail = b_smem + (size_t)b_slots * B_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* a_full = bars;                 // [4]  leader's copy collects the bytes of both CTAs
  uint64_t* a_empty = bars + 4;            // [4]  local, multicast commit
  uint64_t* b_full = bars + 8;             // [8]  leader's copy (ring) / [0] = weights loaded (resident)
  uint64_t* b_empty = bars + 16;           // [8]  local, multicast commit
  uint64_t* tmem_full = bars + 24;         // [4]  local, multicast commit
  uint64_t* tmem_empty = bars + 28;        // [4]  leader's copy: 4 epilogue warps of each CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  float* bias_s = reinterpret_cast<float*>(tail + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid with cluster dims (2,1,1): %cluster_ctarank == blockIdx.x & 1; derived from blockIdx so that the
  // compiler knows it is warp-uniform (keeps the MMA / TMA issue loops on the uniform datapath)
  const uint32_t rank = blockIdx.x & 1;
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  for (int i = threadIdx.x; i < p.cout; i += NUM_THREADS) bias_s[i] = p.bias[i];
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out);
    if (SPLIT) {
      prefetch_tmap(&tmap_a_lo);
      prefetch_tmap(&tmap_b_lo);
      prefetch_tmap(&tmap_out_lo);
    }
    for (int s = 0; s < 4; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 8; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 4; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(ACC * BLOCK_N))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // barrier inits + TMEM allocation of both CTAs visible cluster-wide
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_bytes = (uint32_t)(HALO_H * p.hbw * 128);

  // work item -> (channel tile, m-tile pair); this CTA's pixel tile is m-tile 2*pair + rank
  auto decode = [&](int item, int& n0, int& img, int& h0, int& w0) {
    const int n_tile = item % p.num_n_tiles;
    int m_tile = (item / p.num_n_tiles) * 2 + (int)rank;
    const int twi = m_tile % p.tiles_w; m_tile /= p.tiles_w;
    const int thi = m_tile % p.tiles_h;
    img = m_tile / p.tiles_h;              // == n for the padding tile of an odd tile count: TMA zero-fills / clips
    h0 = thi * TH;
    w0 = twi * TW;
    n0 = n_tile * BLOCK_N;
  };

  if (warp == 0) {
    // =============================== TMA producer (both CTAs; bytes are counted on the leader's barriers) ===============
    if (RESIDENT) {
      if (elect_one()) {
        const int wn0 = (cluster_id % p.num_n_tiles) * BLOCK_N + (int)rank * HALF_N;
        const uint32_t bar = map_to_rank(&b_full[0], 0);
        if (leader) mbar_expect_tx(&b_full[0], (uint32_t)(2 * chunks * 9 * B_TILE));
        for (int cb = 0; cb < chunks; ++cb)
          for (int tap = 0; tap < 9; ++tap)
            tma_load_2d_pair(b_smem + (size_t)(cb * 9 + tap) * B_TILE, &tmap_b, bar, tap * p.cin + cb * BLOCK_K, wn0);
      }
      __syncwarp();
    }
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int item = cluster_id; item < p.num_items; item += num_clusters) {
      int n0, img, h0, w0;
      decode(item, n0, img, h0, w0);
      for (int cb = 0; cb < chunks; ++cb) {
        mbar_wait(&a_empty[as], aph ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&a_full[as], 2 * NSPLIT * a_bytes);
          tma_load_4d_pair(a_smem + (size_t)as * a_stride, &tmap_a, map_to_rank(&a_full[as], 0), cb * BLOCK_K, w0 - 1,
                           h0 - 1, img);
          if (SPLIT)
            tma_load_4d_pair(a_smem + (size_t)as * a_stride + p.a_stage_bytes, &tmap_a_lo, map_to_rank(&a_full[as], 0),
                             cb * BLOCK_K, w0 - 1, h0 - 1, img);
        }
        __syncwarp();
        if (++as == a_stages) { as = 0; aph ^= 1; }
        if (!RESIDENT) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&b_full[bs], 2 * B_SLOT);
              tma_load_2d_pair(b_smem + (size_t)bs * B_SLOT, &tmap_b, map_to_rank(&b_full[bs], 0), tap * p.cin + cb * BLOCK_K,
                               n0 + (int)rank * HALF_N);
              if (SPLIT)
                tma_load_2d_pair(b_smem + (size_t)bs * B_SLOT + B_TILE, &tmap_b_lo, map_to_rank(&b_full[bs], 0),
                                 tap * p.cin + cb * BLOCK_K, n0 + (int)rank * HALF_N);
            }
            __syncwarp();
            if (++bs == b_stages) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (leader) {
      constexpr uint32_t idesc = make_idesc(256, BLOCK_N);
      const uint32_t a_sbo = (uint32_t)(p.hbw * 128);
      if (RESIDENT) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int local = 0;
      for (int item = cluster_id; item < p.num_items; item += num_clusters, ++local) {
        const int acc = local & (ACC - 1);
        const uint32_t acc_phase = (local >> ACC_LOG2) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int cb = 0; cb < chunks; ++cb) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint64_t da0 = make_desc_sw128(smem_u32(a_smem + (size_t)as * a_stride), a_sbo, 0);
          const uint64_t da0_lo = SPLIT ? make_desc_sw128(smem_u32(a_smem + (size_t)as * a_stride + p.a_stage_bytes), a_sbo, 0) : da0;
          if (RESIDENT) {
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int r = tap / 3, s = tap - r * 3;
                const uint64_t da_tap = da0 + (uint64_t)((r * p.hbw + s) * 8);
                const uint64_t db_tap = make_desc_sw128(smem_u32(b_smem + (size_t)(cb * 9 + tap) * B_TILE), 1024, 0);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                  umma_bf16_pair(d_tmem, da_tap + (uint64_t)(k * 2), db_tap + (uint64_t)(k * 2), idesc,
                                 (cb > 0 || tap > 0 || k > 0) ? 1u : 0u);
              }
              umma_commit_pair(&a_empty[as]);
              if (cb == chunks - 1) umma_commit_pair(&tmem_full[acc]);
            }
            __syncwarp();
          } else {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              if (elect_one()) {
                const int r = tap / 3, s = tap - r * 3;
                const uint64_t da_tap = da0 + (uint64_t)((r * p.hbw + s) * 8);
                const uint64_t db_tap = make_desc_sw128(smem_u32(b_smem + (size_t)bs * B_SLOT), 1024, 0);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                  umma_bf16_pair(d_tmem, da_tap + (uint64_t)(k * 2), db_tap + (uint64_t)(k * 2), idesc,
                                 (cb > 0 || tap > 0 || k > 0) ? 1u : 0u);
                  if (SPLIT) {   // + hi * lo + lo * hi (lo * lo is below fp32 resolution)
                    const uint64_t da_lo = da0_lo + (uint64_t)((r * p.hbw + s) * 8);
                    const uint64_t db_lo = make_desc_sw128(smem_u32(b_smem + (size_t)bs * B_SLOT + B_TILE), 1024, 0);
                    umma_bf16_pair(d_tmem, da_tap + (uint64_t)(k * 2), db_lo + (uint64_t)(k * 2), idesc, 1u);
                    umma_bf16_pair(d_tmem, da_lo + (uint64_t)(k * 2), db_tap + (uint64_t)(k * 2), idesc, 1u);
                  }
                }
                umma_commit_pair(&b_empty[bs]);
                if (tap == 8) {
                  umma_commit_pair(&a_empty[as]);
                  if (cb == chunks - 1) umma_commit_pair(&tmem_full[acc]);
                }
              }
              __syncwarp();
              if (++bs == b_stages) { bs = 0; bph ^= 1; }
            }
          }
          if (++as == a_stages) { as = 0; aph ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue (both CTAs, own 128 TMEM lanes / own pixel tile) ===============================
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const bool elected = warp == 2 && lane == 0;
    int local = 0, sub = 0;
    for (int item = cluster_id; item < p.num_items; item += num_clusters, ++local) {
      const int acc = local & (ACC - 1);
      const uint32_t acc_phase = (local >> ACC_LOG2) & 1;
      int n0, img, h0, w0;
      decode(item, n0, img, h0, w0);
      const int py = h0 + (m >> 3), px = w0 + (m & 7);
      const bool pix_ok = img < p.n && py < p.h && px < p.w;
      const size_t pix_off = (((size_t)img * p.h + py) * p.w + px) * p.cout;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c64 = 0; c64 < BLOCK_N / 64; ++c64, ++sub) {
        // bf16: two staging tiles used alternately; SPLIT: tile 0 = hi, tile 1 = lo of the same sub-tile
        uint8_t* stg = stg_smem + (SPLIT ? 0 : (sub & 1)) * (128 * 128);
        if (elected) {
          if (SPLIT) tma_store_wait_read<0>();
          else tma_store_wait_read<1>();
        }
        named_bar_sync(2, 128);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + c64 * 64 + half * 32, r);
          tmem_ld_wait();
          uint4 mk[4];
          if (p.epi == kEpiMask || p.epi == kEpiDivide) {
#pragma unroll
            for (int q = 0; q < 4; ++q) mk[q] = make_uint4(0u, 0u, 0u, 0u);
            if (pix_ok) {
              const uint4* mp = reinterpret_cast<const uint4*>(p.mask + pix_off + n0 + c64 * 64 + half * 32);
#pragma unroll
              for (int q = 0; q < 4; ++q) mk[q] = __ldg(mp + q);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
            if (p.epi == kEpiBiasRelu) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                v[j] = fmaxf(__uint_as_float(r[q * 8 + j]) + bias_s[n0 + c64 * 64 + half * 32 + q * 8 + j], 0.f);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[q * 8 + j]);
              const uint32_t mw[4] = {mk[q].x, mk[q].y, mk[q].z, mk[q].w};
              if (p.epi == kEpiMask) {      // ReLU of the layer below: pass the gradient where its output was > 0
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if ((mw[j] & 0x00007fffu) == 0u) v[2 * j] = 0.f;
                  if ((mw[j] & 0x7fff0000u) == 0u) v[2 * j + 1] = 0.f;
                }
              } else if (p.epi == kEpiDivide) {   // gradient w.r.t. a pooled tensor, handed on as grad / pooled
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float d0 = __uint_as_float(mw[j] << 16), d1 = __uint_as_float(mw[j] & 0xffff0000u);
                  v[2 * j] = pix_ok ? v[2 * j] / d0 : 0.f;
                  v[2 * j + 1] = pix_ok ? v[2 * j + 1] / d1 : 0.f;
                }
              }
            }
            uint4 t;
            t.x = Vec<bf16>::pack2(v[0], v[1]);
            t.y = Vec<bf16>::pack2(v[2], v[3]);
            t.z = Vec<bf16>::pack2(v[4], v[5]);
            t.w = Vec<bf16>::pack2(v[6], v[7]);
            if (SPLIT) {
              *reinterpret_cast<uint4*>(stg + m * 128 + (((half * 4 + q) ^ (m & 7)) << 4)) = t;
              float l[8];
              const uint32_t hw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                l[2 * j] = v[2 * j] - __uint_as_float(hw[j] << 16);
                l[2 * j + 1] = v[2 * j + 1] - __uint_as_float(hw[j] & 0xffff0000u);
              }
              uint4 tl;
              tl.x = Vec<bf16>::pack2(l[0], l[1]);
              tl.y = Vec<bf16>::pack2(l[2], l[3]);
              tl.z = Vec<bf16>::pack2(l[4], l[5]);
              tl.w = Vec<bf16>::pack2(l[6], l[7]);
              *reinterpret_cast<uint4*>(stg + 128 * 128 + m * 128 + (((half * 4 + q) ^ (m & 7)) << 4)) = tl;
            } else if (p.strips) {
              // [channel][pixel]: a warp's 32 pixels of one channel are 64 contiguous bytes (conflict-free 2-byte stores)
              unsigned short* col = reinterpret_cast<unsigned short*>(stg) + (half * 32 + q * 8) * 128 + m;
              const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                col[(2 * j) * 128] = (unsigned short)(tw[j] & 0xffffu);
                col[(2 * j + 1) * 128] = (unsigned short)(tw[j] >> 16);
              }
            } else {
              *reinterpret_cast<uint4*>(stg + m * 128 + (((half * 4 + q) ^ (m & 7)) << 4)) = t;
            }
          }
        }
        if (c64 == BLOCK_N / 64 - 1) {                // accumulator fully read: hand it back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(map_to_rank(&tmem_empty[acc], 0));
        }
        fence_proxy_async_smem();
        named_bar_sync(3, 128);
        if (elected) {
          if (p.strips) tma_store_4d(&tmap_out, stg, h0 * 8, w0 >> 3, n0 + c64 * 64, img);
          else tma_store_4d(&tmap_out, stg, n0 + c64 * 64, w0, h0, img);
          if (SPLIT) tma_store_4d(&tmap_out_lo, stg + 128 * 128, n0 + c64 * 64, w0, h0, img);
          tma_store_commit();
        }
      }
    }
    if (elected) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer may still signal barriers in this CTA's shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(ACC * BLOCK_N))
                 : "memory");
  }
}

template <int BLOCK_N, bool RESIDENT, bool SPLIT>
static int launch_impl(const ConvTcPlan& plan, int sm_count, cudaStream_t st) {
  constexpr int NSPLIT = SPLIT ? 2 : 1;
  constexpr int B_TILE = (BLOCK_N / 2) * BLOCK_K * 2;
  Params p;
  p.n = plan.n; p.h = plan.h; p.w = plan.w; p.cin = plan.cin; p.cout = plan.cout;
  p.tiles_w = plan.tiles_w; p.tiles_h = plan.tiles_h;
  p.num_n_tiles = plan.cout / BLOCK_N;
  const int m_tiles = plan.n * plan.tiles_h * plan.tiles_w;
  p.num_m_pairs = (m_tiles + 1) / 2;
  p.num_items = p.num_m_pairs * p.num_n_tiles;
  p.hbw = plan.halo_hbw;
  p.a_stage_bytes = (HALO_H * plan.halo_hbw * 128 + 1023) & ~1023;
  p.bias = plan.bias; p.epi = plan.epi; p.mask = plan.mask;
  p.strips = plan.strips_out ? 1 : 0;
  if (p.strips && p.epi != kEpiBiasRelu) return -1;
  if (SPLIT && (p.epi != kEpiBiasRelu || p.strips)) return -1;   // bf16x3: forward convolutions only
  const int chunks = plan.cin / BLOCK_K;
  const int budget = 227 * 1024 - 1024 /*align*/ - 512 /*barriers*/ - 2048 /*bias*/ - STG_BYTES;
  int a_stages, b_stages = 0;
  if (RESIDENT) {
    const int wbytes = chunks * 9 * B_TILE;
    a_stages = (budget - wbytes) / p.a_stage_bytes;
    if (a_stages > 4) a_stages = 4;
    if (a_stages < 2) return -1;
  } else {
    a_stages = 2;
    b_stages = (budget - a_stages * NSPLIT * p.a_stage_bytes) / (NSPLIT * B_TILE);
    if (b_stages > 8) b_stages = 8;
    if (b_stages < 3) return -1;
  }
  const int smem = STG_BYTES + a_stages * NSPLIT * p.a_stage_bytes + (RESIDENT ? chunks * 9 : b_stages) * NSPLIT * B_TILE + 1024 +
                   512 + 2048;
  ensure_dynamic_smem(conv3x3_tc_pair_kernel<BLOCK_N, RESIDENT, SPLIT>, smem);
  int clusters = sm_count / 2;
  if (p.num_items < clusters) clusters = p.num_items;
  if (RESIDENT && clusters > p.num_n_tiles) clusters -= clusters % p.num_n_tiles;
  conv3x3_tc_pair_kernel<BLOCK_N, RESIDENT, SPLIT><<<2 * clusters, NUM_THREADS, smem, st>>>(
      plan.tmap_a, plan.tmap_b, plan.tmap_out, plan.tmap_a_lo, plan.tmap_b_lo, plan.tmap_out_lo, p, a_stages, b_stages);
  return 1;
}

}  // namespace tcp

// Returns launches (1) or -1 if the configuration does not fit.  plan.tmap_b must have been encoded with a box of
// block_n / 2 rows (conv_tc_make_plan does that when plan.pair is set).
int launch_conv3x3_tc_pair(const ConvTcPlan& plan, int sm_count, cudaStream_t st) {
  if (plan.split) {
    if (plan.block_n == 128) return tcp::launch_impl<128, false, true>(plan, sm_count, st);
    if (plan.block_n == 256) return tcp::launch_impl<256, false, true>(plan, sm_count, st);
    return -1;
  }
  const bool resident = (size_t)9 * plan.cin * (plan.block_n / 2) * 2 <= 150 * 1024;
  if (plan.block_n == 128)
    return resident ? tcp::launch_impl<128, true, false>(plan, sm_count, st) : tcp::launch_impl<128, false, false>(plan, sm_count, st);
  if (plan.block_n == 256)
    return resident ? tcp::launch_impl<256, true, false>(plan, sm_count, st) : tcp::launch_impl<256, false, false>(plan, sm_count, st);
  return -1;
}

}  // namespace nqa
