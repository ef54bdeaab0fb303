// Halo-reuse variant of the tcgen05 3x3 convolution for the thin layers (Cin <= 128), where the plain
// kernel (conv_tc.cu) is bound by re-fetching the A tile nine times from L2.
//
//   * output tile = 16 rows x 8 columns of pixels (M = 128, m = 8*row + col);
//   * per 64-channel chunk ONE 4-D TMA box {64 ch, HBW, 18, 1} brings the (16+2) x HBW halo (HBW >= 10) into
//     shared memory, one 128-byte SWIZZLE_128B row per pixel, zero-filled outside the image;
//   * the A operand of tap (r, s) is the SAME buffer viewed through a shifted UMMA descriptor: start address
//     + (r*HBW + s)*128 bytes, 8-row groups (= one tile row) HBW*128 bytes apart (SBO);
//   * weights either stay resident in shared memory for the whole kernel (9*Cin*BLOCK_N*2 bytes; conv1_2 and
//     conv2_1) or stream through their own mbarrier ring (conv2_2, conv3_1);
//   * everything else (persistent tile loop, TMEM double buffering, epilogue) as in conv_tc.cu.
//
// The 128-byte swizzle is a function of shared-memory address bits; TMA writes and UMMA reads agree as
// long as the descriptor's base-offset field describes the phase of the (non 1024-byte-aligned) start
// address.  `desc_mode` selects how that field is filled (validated on hardware by tests/test_gpu_ops.py).
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace nqa {
namespace tch {

using namespace tc;

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int TH = 16, TW = 8, HALO_H = TH + 2;
constexpr int NUM_THREADS = 192;
constexpr int STG_BYTES = 2 * 128 * 128;   // two 64-channel sub-tiles of 128 pixels

struct Params {
  int n, h, w, cin, cout;
  int tiles_w, tiles_h, num_n_tiles, num_tiles;
  int hbw;          // halo box width in pixels (10 or 16)
  int desc_mode;    // 0: base_offset = 0, 1: base_offset = (start >> 7) & 7
  int a_stage_bytes;
  const float* bias;
  bf16* out;
  int epi;             // ConvEpilogue
  const bf16* mask;    // kEpiMask: saved activation of the layer below, [n,h,w,cout]
  float* fg_x;         // kEpiFrameGrad
  float* fg_y;
  int fg_nx;
};

template <int BLOCK_N, bool RESIDENT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_tc_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ CUtensorMap tmap_out, const Params p, int a_stages, int b_stages) {
  constexpr int B_TILE = BLOCK_N * BLOCK_K * 2;
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET from the __shared__ array (a round trip through uintptr_t would turn every access into a generic
  // LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  const int chunks = p.cin / BLOCK_K;
  uint8_t* stg_smem = smem;                                      // [2][128 x 128 B] output staging for the TMA store
  uint8_t* a_smem = smem + STG_BYTES;                            // [a_stages][a_stage_bytes]
  uint8_t* b_smem = a_smem + a_stages * p.a_stage_bytes;         // resident: [chunks*9][B_TILE]; ring: [b_stages][B_TILE]
  const int b_slots = RESIDENT ? chunks * 9 : b_stages;
  uint8_t* tail = b_smem + (size_t)b_slots * B_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* a_full = bars;                 // [4]
  uint64_t* a_empty = bars + 4;            // [4]
  uint64_t* b_full = bars + 8;             // [8] (ring) / [0] = weights loaded (resident)
  uint64_t* b_empty = bars + 16;           // [8]
  uint64_t* tmem_full = bars + 24;         // [2]
  uint64_t* tmem_empty = bars + 26;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  float* bias_s = reinterpret_cast<float*>(tail + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.cout; i += NUM_THREADS) bias_s[i] = p.bias[i];
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out);
    for (int s = 0; s < 4; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 8; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(2 * BLOCK_N))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_bytes = (uint32_t)(HALO_H * p.hbw * 128);

  if (warp == 0) {
    // =============================== TMA producer (whole warp, one elected lane issues) ===============================
    if (RESIDENT) {
      // all weights of this CTA's channel tile, once: the grid is a multiple of num_n_tiles, so with the
      // channel tile fastest in the tile order a CTA sees the same channel tile for its whole life
      if (elect_one()) {
        const int wn0 = (blockIdx.x % p.num_n_tiles) * BLOCK_N;
        mbar_expect_tx(&b_full[0], (uint32_t)(chunks * 9 * B_TILE));
        for (int cb = 0; cb < chunks; ++cb)
          for (int tap = 0; tap < 9; ++tap)
            tma_load_2d(b_smem + (size_t)(cb * 9 + tap) * B_TILE, &tmap_b, &b_full[0], tap * p.cin + cb * BLOCK_K, wn0);
      }
      __syncwarp();
    }
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.num_n_tiles;
      int m_tile = tile / p.num_n_tiles;
      const int twi = m_tile % p.tiles_w; m_tile /= p.tiles_w;
      const int thi = m_tile % p.tiles_h;
      const int img = m_tile / p.tiles_h;
      const int h0 = thi * TH, w0 = twi * TW, n0 = n_tile * BLOCK_N;
      for (int cb = 0; cb < chunks; ++cb) {
        mbar_wait(&a_empty[as], aph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&a_full[as], a_bytes);
          tma_load_4d(a_smem + (size_t)as * p.a_stage_bytes, &tmap_a, &a_full[as], cb * BLOCK_K, w0 - 1, h0 - 1, img);
        }
        __syncwarp();
        if (++as == a_stages) { as = 0; aph ^= 1; }
        if (!RESIDENT) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&b_full[bs], B_TILE);
              tma_load_2d(b_smem + (size_t)bs * B_TILE, &tmap_b, &b_full[bs], tap * p.cin + cb * BLOCK_K, n0);
            }
            __syncwarp();
            if (++bs == b_stages) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp, one elected lane issues) ===============================
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
    const uint32_t a_sbo = (uint32_t)(p.hbw * 128);
    if (RESIDENT) {
      mbar_wait(&b_full[0], 0);
      tc_fence_after();
    }
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int cb = 0; cb < chunks; ++cb) {
        mbar_wait(&a_full[as], aph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(a_smem + (size_t)as * p.a_stage_bytes);
        // descriptor of tap (0,0), k = 0; taps and k-steps only move the 16-byte-granular start-address field
        const uint64_t da0 = make_desc_sw128(a_base, a_sbo, 0);
        if (RESIDENT) {
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int r = tap / 3, s = tap - r * 3;
              const uint64_t da_tap = da0 + (uint64_t)((r * p.hbw + s) * 8);
              const uint64_t db_tap = make_desc_sw128(smem_u32(b_smem + (size_t)(cb * 9 + tap) * B_TILE), 1024, 0);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                umma_bf16(d_tmem, da_tap + (uint64_t)(k * 2), db_tap + (uint64_t)(k * 2), idesc,
                          (cb > 0 || tap > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&a_empty[as]);
            if (cb == chunks - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
        } else {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            if (elect_one()) {
              const int r = tap / 3, s = tap - r * 3;
              const uint64_t da_tap = da0 + (uint64_t)((r * p.hbw + s) * 8);
              const uint64_t db_tap = make_desc_sw128(smem_u32(b_smem + (size_t)bs * B_TILE), 1024, 0);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                umma_bf16(d_tmem, da_tap + (uint64_t)(k * 2), db_tap + (uint64_t)(k * 2), idesc,
                          (cb > 0 || tap > 0 || k > 0) ? 1u : 0u);
              umma_commit(&b_empty[bs]);
              if (tap == 8) {
                umma_commit(&a_empty[as]);
                if (cb == chunks - 1) umma_commit(&tmem_full[acc]);
              }
            }
            __syncwarp();
            if (++bs == b_stages) { bs = 0; bph ^= 1; }
          }
        }
        if (++as == a_stages) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // =============================== epilogue ===============================
    // thread = pixel (TMEM lane); per 64-channel sub-tile its 128-byte row is staged in the swizzled layout of the
    // output tensor map and ONE TMA store writes the 16 x 8 x 64 box (full lines, edges clipped by the TMA unit).
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const bool elected = warp == 2 && lane == 0;
    int local = 0, sub = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int n_tile = tile % p.num_n_tiles;
      int m_tile = tile / p.num_n_tiles;
      const int twi = m_tile % p.tiles_w; m_tile /= p.tiles_w;
      const int thi = m_tile % p.tiles_h;
      const int img = m_tile / p.tiles_h;
      const int n0 = n_tile * BLOCK_N;
      const int py = thi * TH + (m >> 3), px = twi * TW + (m & 7);
      const bool pix_ok = py < p.h && px < p.w;
      const size_t pix_off = (((size_t)img * p.h + py) * p.w + px) * p.cout;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (p.epi == kEpiFrameGrad) {                   // three fp32 planes per image; no staging, no TMA store
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (pix_ok) {
          const size_t hw = (size_t)p.h * p.w;
          float* dst = img < p.fg_nx ? p.fg_x + (size_t)img * 3 * hw : p.fg_y + (size_t)(img - p.fg_nx) * 3 * hw;
          const size_t o = (size_t)py * p.w + px;
          dst[o] = __uint_as_float(r[0]) / 0.229f;
          dst[hw + o] = __uint_as_float(r[1]) / 0.224f;
          dst[2 * hw + o] = __uint_as_float(r[2]) / 0.225f;
        }
        continue;
      }
#pragma unroll 1
      for (int c64 = 0; c64 < BLOCK_N / 64; ++c64, ++sub) {
        uint8_t* stg = stg_smem + (sub & 1) * (128 * 128);
        if (elected) tma_store_wait_read<1>();        // the store issued two sub-tiles ago has read this buffer
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
            *reinterpret_cast<uint4*>(stg + m * 128 + (((half * 4 + q) ^ (m & 7)) << 4)) = t;
          }
        }
        if (c64 == BLOCK_N / 64 - 1) {                // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        fence_proxy_async_smem();
        named_bar_sync(3, 128);
        if (elected) {
          tma_store_4d(&tmap_out, stg, n0 + c64 * 64, twi * TW, thi * TH, img);
          tma_store_commit();
        }
      }
    }
    if (elected) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BLOCK_N))
                 : "memory");
  }
}

template <int BLOCK_N, bool RESIDENT>
static int launch_impl(const ConvTcPlan& plan, int sm_count, cudaStream_t st) {
  constexpr int B_TILE = BLOCK_N * BLOCK_K * 2;
  Params p;
  p.n = plan.n; p.h = plan.h; p.w = plan.w; p.cin = plan.cin; p.cout = plan.cout;
  p.tiles_w = plan.tiles_w; p.tiles_h = plan.tiles_h;
  p.num_n_tiles = plan.cout / BLOCK_N;
  p.num_tiles = plan.n * plan.tiles_h * plan.tiles_w * p.num_n_tiles;
  p.hbw = plan.halo_hbw; p.desc_mode = plan.halo_desc_mode;
  p.a_stage_bytes = (HALO_H * plan.halo_hbw * 128 + 1023) & ~1023;
  p.bias = plan.bias; p.out = plan.out; p.epi = plan.epi; p.mask = plan.mask;
  p.fg_x = plan.fg_x; p.fg_y = plan.fg_y; p.fg_nx = plan.fg_nx;
  if (plan.epi == kEpiFrameGrad && (BLOCK_N != 64 || plan.cout != 64)) return -1;
  const int chunks = plan.cin / BLOCK_K;
  const int budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - 2048 /*bias*/ - STG_BYTES;
  int a_stages, b_stages = 0;
  if (RESIDENT) {
    const int wbytes = chunks * 9 * B_TILE;
    a_stages = (budget - wbytes) / p.a_stage_bytes;
    if (a_stages > 4) a_stages = 4;
    if (a_stages < 2) return -1;
  } else {
    a_stages = 2;
    b_stages = (budget - a_stages * p.a_stage_bytes) / B_TILE;
    if (b_stages > 8) b_stages = 8;
    if (b_stages < 3) return -1;
  }
  const int smem = STG_BYTES + a_stages * p.a_stage_bytes + (RESIDENT ? chunks * 9 : b_stages) * B_TILE + 1024 + 256 + 2048;
  ensure_dynamic_smem(conv3x3_tc_halo_kernel<BLOCK_N, RESIDENT>, smem);
  int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
  if (RESIDENT) grid -= grid % p.num_n_tiles;
  conv3x3_tc_halo_kernel<BLOCK_N, RESIDENT><<<grid, NUM_THREADS, smem, st>>>(plan.tmap_a, plan.tmap_b, plan.tmap_out, p, a_stages,
                                                                             b_stages);
  return 1;
}

}  // namespace tch

// Returns launches (1) or -1 if the configuration does not fit.
int launch_conv3x3_tc_halo(const ConvTcPlan& plan, int sm_count, cudaStream_t st) {
  const bool resident = (size_t)9 * plan.cin * plan.block_n * 2 <= 150 * 1024;
  if (plan.block_n == 64) return resident ? tch::launch_impl<64, true>(plan, sm_count, st) : tch::launch_impl<64, false>(plan, sm_count, st);
  if (plan.block_n == 128) return resident ? tch::launch_impl<128, true>(plan, sm_count, st) : tch::launch_impl<128, false>(plan, sm_count, st);
  return tch::launch_impl<256, false>(plan, sm_count, st);
}

}  // namespace nqa
