// 3x3 convolution + bias + ReLU as an implicit GEMM on the Blackwell tensor pipe (sm_100a).
//
//   D[pixels, cout] = sum_{tap, cin} A[pixel shifted by tap, cin] * W[cout, tap, cin]
//
//   * activations NHWC bf16; one CTA tile = 128 output pixels arranged as a TH x TW rectangle
//     (TH*TW = 128) x BLOCK_N output channels;
//   * per k-block (one filter tap x 64 input channels) ONE 4-D TMA box load {64 ch, TW, TH, 1} at the
//     tap-shifted coordinate brings the A tile into shared memory in the canonical K-major
//     SWIZZLE_128B layout (out-of-image rows/columns are zero-filled by TMA = the conv's zero
//     padding); one 2-D TMA box {64 k, BLOCK_N} brings the weight tile;
//   * tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) issued by one thread, accumulators in TMEM,
//     double-buffered (2 x BLOCK_N columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * epilogue warps: tcgen05.ld -> +bias -> ReLU -> bf16 -> 16-byte global stores (NHWC);
//   * persistent: grid = #SMs, static round-robin over (pixel tile, channel tile) with the channel
//     tile fastest so CTAs that share an A tile run back to back and hit L2;
//   * bf16x3 ("split") mode: activations and weights are carried as bf16 hi+lo pairs and each k-step
//     issues hi*hi + hi*lo + lo*hi into the same fp32 accumulator (fp32-class accuracy).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_idx % 4).
//
// Reference semantics: torchvision vgg16.features convs as used by
// nerf_qa/DISTS_pytorch/DISTS_pt.py:36-49 and nerf_qa/ADISTS/ADISTS.py:44-57.
#include <stdio.h>
#include <stdlib.h>

#ifndef NQA_CONV_HALO_DEFAULT
#define NQA_CONV_HALO_DEFAULT 1
#endif

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace nqa {

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                            // bf16 elements = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;    // 16 KB
constexpr int NUM_THREADS = 192;
struct Params {
  int n, h, w, cin, cout;
  int tw_log2, tiles_w, tiles_h;
  int num_n_tiles, num_tiles;
  const float* bias;
  bf16* out;
  bf16* out_lo;
  int epi;              // ConvEpilogue: bias + ReLU (forward) or the data-gradient epilogues (bf16x3 backward)
  const bf16* mask;     // kEpiMask / kEpiDivide: saved activation of the output's shape (hi carrier)
  const bf16* mask_lo;  // kEpiDivide with a bf16x3 carrier: residual of the saved activation
};

template <int BLOCK_N, bool SPLIT>
struct Cfg {
  static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = (A_TILE_BYTES + B_TILE_BYTES) * (SPLIT ? 2 : 1);
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // 128 / 256 / 512: powers of two >= 32
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2048 /*bias*/;
};

template <int BLOCK_N, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_lo,
                  const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b_lo,
                  const Params p) {
  using C = Cfg<BLOCK_N, SPLIT>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET from the __shared__ array (a round trip through uintptr_t would turn every access into a generic
  // LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;    // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES + 256);  // [cout <= 512]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = 1 << p.tw_log2, th = BLOCK_M >> p.tw_log2;
  const int kb_per_tile = 9 * (p.cin / BLOCK_K);

  for (int i = threadIdx.x; i < p.cout; i += NUM_THREADS) bias_s[i] = p.bias[i];
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (SPLIT) {
      prefetch_tmap(&tmap_a_lo);
      prefetch_tmap(&tmap_b_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (whole warp, one elected lane issues) ===============================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.num_n_tiles;
      int m_tile = tile / p.num_n_tiles;
      const int twi = m_tile % p.tiles_w; m_tile /= p.tiles_w;
      const int thi = m_tile % p.tiles_h;
      const int img = m_tile / p.tiles_h;
      const int h0 = thi * th, w0 = twi * tw, n0 = n_tile * BLOCK_N;
      for (int cb = 0; cb < p.cin / BLOCK_K; ++cb) {
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            uint8_t* sb = sa + A_TILE_BYTES;
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            const int r = tap / 3, s = tap % 3;
            tma_load_4d(sa, &tmap_a, &full_bar[stage], cb * BLOCK_K, w0 + s - 1, h0 + r - 1, img);
            tma_load_2d(sb, &tmap_b, &full_bar[stage], tap * p.cin + cb * BLOCK_K, n0);
            if (SPLIT) {
              uint8_t* sa_lo = sb + C::B_TILE_BYTES;
              uint8_t* sb_lo = sa_lo + A_TILE_BYTES;
              tma_load_4d(sa_lo, &tmap_a_lo, &full_bar[stage], cb * BLOCK_K, w0 + s - 1, h0 + r - 1, img);
              tma_load_2d(sb_lo, &tmap_b_lo, &full_bar[stage], tap * p.cin + cb * BLOCK_K, n0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp, one elected lane issues) ===============================
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < kb_per_tile; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + A_TILE_BYTES;
          const uint64_t da0 = make_smem_desc(sa), db0 = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advancing K by 16 elements = 32 bytes = 2 units of the descriptor's 16-byte address field
            const uint64_t da = da0 + (uint64_t)(k * 2), db = db0 + (uint64_t)(k * 2);
            umma_bf16(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (SPLIT) {
              const uint64_t da_lo = make_smem_desc(sb + C::B_TILE_BYTES) + (uint64_t)(k * 2);
              const uint64_t db_lo = make_smem_desc(sb + C::B_TILE_BYTES + A_TILE_BYTES) + (uint64_t)(k * 2);
              umma_bf16(d_tmem, da, db_lo, idesc, 1u);
              umma_bf16(d_tmem, da_lo, db, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);                       // frees the smem slot when the MMAs retire
          if (kb == kb_per_tile - 1) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    const int m = quarter * 32 + lane;
    const int lh = m >> p.tw_log2, lw = m & (tw - 1);
    int local = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int n_tile = tile % p.num_n_tiles;
      int m_tile = tile / p.num_n_tiles;
      const int twi = m_tile % p.tiles_w; m_tile /= p.tiles_w;
      const int thi = m_tile % p.tiles_h;
      const int img = m_tile / p.tiles_h;
      const int oy = thi * th + lh, ox = twi * tw + lw, n0 = n_tile * BLOCK_N;
      const bool valid = oy < p.h && ox < p.w;
      const size_t obase = (((size_t)img * p.h + oy) * p.w + ox) * p.cout + n0;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + ch * 32, r);
        tmem_ld_wait();
        float v[32];
        if (p.epi == kEpiBiasRelu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r[j]) + bias_s[n0 + ch * 32 + j], 0.f);
        } else {
          // data-gradient epilogues (kernels.h ConvEpilogue): plain sum, sum masked by the ReLU of the layer below, or
          // sum / pooled activation; the saved tensor has the output's shape
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.epi != kEpiPlain && valid) {
            const uint4* mh = reinterpret_cast<const uint4*>(p.mask + obase + ch * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float m[8];
              Vec<bf16>::unpack(__ldg(mh + q), m);
              if (p.epi == kEpiDivide && p.mask_lo) {
                float ml[8];
                Vec<bf16>::unpack(__ldg(reinterpret_cast<const uint4*>(p.mask_lo + obase + ch * 32) + q), ml);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] += ml[j];
              }
#pragma unroll
              for (int j = 0; j < 8; ++j)
                v[q * 8 + j] = p.epi == kEpiMask ? (m[j] > 0.f ? v[q * 8 + j] : 0.f) : v[q * 8 + j] / m[j];
            }
          }
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + obase + ch * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 t;
            t.x = Vec<bf16>::pack2(v[q * 8 + 0], v[q * 8 + 1]);
            t.y = Vec<bf16>::pack2(v[q * 8 + 2], v[q * 8 + 3]);
            t.z = Vec<bf16>::pack2(v[q * 8 + 4], v[q * 8 + 5]);
            t.w = Vec<bf16>::pack2(v[q * 8 + 6], v[q * 8 + 7]);
            dst[q] = t;
          }
          if (SPLIT) {
            uint4* dlo = reinterpret_cast<uint4*>(p.out_lo + obase + ch * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float l[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) l[j] = v[q * 8 + j] - __bfloat162float(__float2bfloat16_rn(v[q * 8 + j]));
              uint4 t;
              t.x = Vec<bf16>::pack2(l[0], l[1]);
              t.y = Vec<bf16>::pack2(l[2], l[3]);
              t.z = Vec<bf16>::pack2(l[4], l[5]);
              t.w = Vec<bf16>::pack2(l[6], l[7]);
              dlo[q] = t;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}


static int choose_tw_log2(int h, int w) {
  // pick the TH x TW = 128 rectangle that wastes the fewest MMA rows on this image size
  int best = 3;
  long best_cost = -1;
  for (int l = 3; l <= 7; ++l) {
    const int tw = 1 << l, th = 128 >> l;
    const long cost = (long)ceil_div(w, tw) * tw * (long)ceil_div(h, th) * th;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = l; }
  }
  return best;
}

static bool encode_act_map(CUtensorMap* map, const bf16* ptr, int n, int h, int w, int c, int tw, int th) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  const cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)tw, (cuuint32_t)th, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool encode_w_map(CUtensorMap* map, const bf16* ptr, int cout, int k, int block_n) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)cout};
  const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)block_n};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BLOCK_N, bool SPLIT>
static int launch_impl(const ConvTcPlan& plan, cudaStream_t st) {
  using C = Cfg<BLOCK_N, SPLIT>;
  ensure_dynamic_smem(conv3x3_tc_kernel<BLOCK_N, SPLIT>, C::SMEM_BYTES);
  Params p;
  p.n = plan.n; p.h = plan.h; p.w = plan.w; p.cin = plan.cin; p.cout = plan.cout;
  p.tw_log2 = plan.tw_log2; p.tiles_w = plan.tiles_w; p.tiles_h = plan.tiles_h;
  p.num_n_tiles = plan.cout / BLOCK_N;
  p.num_tiles = plan.n * plan.tiles_h * plan.tiles_w * p.num_n_tiles;
  p.bias = plan.bias; p.out = plan.out; p.out_lo = plan.out_lo;
  p.epi = plan.epi; p.mask = plan.mask; p.mask_lo = plan.mask_lo;
  const int grid = p.num_tiles < plan.sm_count ? p.num_tiles : plan.sm_count;
  conv3x3_tc_kernel<BLOCK_N, SPLIT><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(plan.tmap_a, plan.tmap_a_lo, plan.tmap_b,
                                                                             plan.tmap_b_lo, p);
  return 1;
}

}  // namespace tc

int conv_tc_make_act_map(CUtensorMap* map, const bf16* ptr, int n, int h, int w, int c, int box_w, int box_h) {
  return tc::encode_act_map(map, ptr, n, h, w, c, box_w, box_h) ? 0 : 1;
}

int conv_tc_make_weight_map(CUtensorMap* map, const bf16* wk, int cout, int k, int block_n) {
  return tc::encode_w_map(map, wk, cout, k, block_n) ? 0 : 1;
}

int make_strip_map(CUtensorMap* map, const bf16* ptr, int planes, int h, int w, int box_elems, int box_strips) {
  tc::EncodeTiledFn fn = tc::get_encode_fn();
  if (!fn || (w & 7) != 0) return 1;
  const cuuint64_t dims[3] = {(cuuint64_t)8 * h, (cuuint64_t)(w / 8), (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)h * 16, (cuuint64_t)h * w * 2};
  const cuuint32_t box[3] = {(cuuint32_t)box_elems, (cuuint32_t)box_strips, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

int make_strip_out_map(CUtensorMap* map, const bf16* ptr, int n, int c, int h, int w) {
  tc::EncodeTiledFn fn = tc::get_encode_fn();
  if (!fn || (w & 7) != 0) return 1;
  const cuuint64_t dims[4] = {(cuuint64_t)8 * h, (cuuint64_t)(w / 8), (cuuint64_t)c, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)h * 16, (cuuint64_t)h * w * 2, (cuuint64_t)c * h * w * 2};
  const cuuint32_t box[4] = {128, 1, 64, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

int make_plane_map_sw64(CUtensorMap* map, const bf16* ptr, int planes, int h, int w, int box_h) {
  tc::EncodeTiledFn fn = tc::get_encode_fn();
  if (!fn) return 1;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)w * 2, (cuuint64_t)h * w * 2};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

// NQA_CONV_HALO: 0 = plain kernel everywhere; 1..4 = halo-reuse kernel with
// (halo width, base-offset mode) = (10,0) (10,1) (16,0) (16,1).
ConvTcConfig conv_tc_config_from_env(int sm_count) {
  ConvTcConfig c;
  c.sm_count = sm_count > 0 ? sm_count : kNumSMs;
  c.halo_mode = NQA_CONV_HALO_DEFAULT;
  if (const char* e = getenv("NQA_CONV_HALO")) c.halo_mode = atoi(e);
  if (c.halo_mode < 0 || c.halo_mode > 4) c.halo_mode = 0;
  if (const char* e = getenv("NQA_CONV_HALO_MAXCIN")) c.halo_max_cin = atoi(e);
  if (const char* e = getenv("NQA_CONV_N64")) c.n64 = atoi(e) != 0;
  if (const char* e = getenv("NQA_CONV_PAIR")) c.pair = atoi(e) != 0;
  if (const char* e = getenv("NQA_PAIR_MINCIN")) c.pair_min_cin = atoi(e);
  if (const char* e = getenv("NQA_PAIR_SPLIT")) c.split_pair = atoi(e) != 0;
  return c;
}

bool conv_tc_uses_pair(const ConvTcConfig& cfg, int cin, int cout, bool split) {
  return cfg.halo_mode > 0 && !split && cin <= cfg.halo_max_cin && cin % tc::BLOCK_K == 0 && cfg.pair && cout >= 128 &&
         cin >= cfg.pair_min_cin && (cfg.sm_count & 1) == 0;
}

int conv_tc_make_plan(ConvTcPlan* plan, const ConvTcConfig& cfg, const bf16* in, const bf16* in_lo, const bf16* wk,
                      const bf16* wk_lo, const float* bias, bf16* out, bf16* out_lo, int n, int h, int w, int cin, int cout,
                      bool split, const char** err, bool strips_out, bool split_on_pairs) {
  if (cin % tc::BLOCK_K != 0 || cout % 64 != 0 || cout > 512) { *err = "conv_tc: cin must be a multiple of 64, cout a multiple of 64 and <= 512"; return 1; }
  plan->n = n; plan->h = h; plan->w = w; plan->cin = cin; plan->cout = cout;
  plan->sm_count = cfg.sm_count;
  plan->split = split; plan->bias = bias; plan->out = out; plan->out_lo = out_lo;
  plan->tw_log2 = tc::choose_tw_log2(h, w);
  int tw = 1 << plan->tw_log2, th = 128 >> plan->tw_log2;
  // bf16x3 stages are twice as large, so cap its channel tile at 128 to keep >= 3 pipeline stages
  plan->block_n = cout >= 256 ? (split ? 128 : 256) : cout >= 128 ? 128 : 64;
  // thin layers: halo-reuse kernel (A tile fetched once instead of nine times)
  plan->epi = kEpiBiasRelu;
  plan->mask = nullptr;
  plan->mask_lo = nullptr;
  plan->fg_x = plan->fg_y = nullptr;
  plan->fg_nx = 0;
  plan->halo = false;
  plan->pair = false;
  plan->strips_out = false;
  plan->halo_hbw = 10;
  plan->halo_desc_mode = 0;
  const int halo_mode = cfg.halo_mode;
  // (bf16x3: conv2_1, 64 -> 128 channels, goes to the pair kernel too -- the alternative is the per-tap kernel, not the
  // single-CTA halo kernel that serves it on the bf16 path)
  const int pair_min_cin = split ? 64 : cfg.pair_min_cin;
  const bool split_pair = split && split_on_pairs && cfg.pair && cfg.split_pair && halo_mode > 0 && cin <= cfg.halo_max_cin &&
                          cout >= 128 && cin >= pair_min_cin && (cfg.sm_count & 1) == 0;
  if (halo_mode > 0 && (!split || split_pair) && cin <= cfg.halo_max_cin) {
    plan->halo = true;
    plan->halo_desc_mode = 0;   // base-offset field 0: validated on hardware (swizzle follows absolute address bits)
    plan->halo_hbw = ((halo_mode - 1) & 2) ? 16 : 10;
    plan->tw_log2 = 3;
    tw = 8;
    th = 16;
    // 128 -> 128 channels: two 64-channel tiles whose weights (147 KB) stay resident in shared memory beat one
    // 128-channel tile whose weights stream from L2 for every pixel tile (NQA_CONV_N64=0 restores the latter)
    if (cfg.n64 && cin == 128 && cout == 128) plan->block_n = 64;
    // CTA pairs (tcgen05 cta_group::2): 256 pixels x block_n channels per instruction, each CTA holds half the weights
    if (cfg.pair && cout >= 128 && cin >= pair_min_cin && (cfg.sm_count & 1) == 0) {
      plan->pair = true;
      plan->block_n = cout >= 256 ? 256 : 128;
    }
  }
  plan->tiles_w = ceil_div(w, tw);
  plan->tiles_h = ceil_div(h, th);
  if (!tc::get_encode_fn()) { *err = "conv_tc: cuTensorMapEncodeTiled entry point not available"; return 1; }
  bool ok = (plan->halo ? tc::encode_act_map(&plan->tmap_a, in, n, h, w, cin, plan->halo_hbw, th + 2)
                        : tc::encode_act_map(&plan->tmap_a, in, n, h, w, cin, tw, th)) &&
            tc::encode_w_map(&plan->tmap_b, wk, cout, 9 * cin, plan->pair ? plan->block_n / 2 : plan->block_n);
  if (strips_out) {
    if (!plan->pair) { *err = "conv_tc: strip-plane output needs the CTA-pair kernel"; return 1; }
    plan->strips_out = true;
    ok = ok && make_strip_out_map(&plan->tmap_out, out, n, cout, h, w) == 0;
  } else if (plan->halo) {
    ok = ok && tc::encode_act_map(&plan->tmap_out, out, n, h, w, cout, tw, th);
  }
  if (split && plan->pair)
    ok = ok && tc::encode_act_map(&plan->tmap_a_lo, in_lo, n, h, w, cin, plan->halo_hbw, th + 2) &&
         tc::encode_w_map(&plan->tmap_b_lo, wk_lo, cout, 9 * cin, plan->block_n / 2) &&
         tc::encode_act_map(&plan->tmap_out_lo, out_lo, n, h, w, cout, tw, th);
  else if (split)
    ok = ok && tc::encode_act_map(&plan->tmap_a_lo, in_lo, n, h, w, cin, tw, th) &&
         tc::encode_w_map(&plan->tmap_b_lo, wk_lo, cout, 9 * cin, plan->block_n);
  else {
    plan->tmap_a_lo = plan->tmap_a;
    plan->tmap_b_lo = plan->tmap_b;
  }
  if (!ok) { *err = "conv_tc: cuTensorMapEncodeTiled failed"; return 1; }
  return 0;
}

int launch_conv3x3_tc(const ConvTcPlan& plan, cudaStream_t st) {
  if (plan.epi == kEpiFrameGrad && !(plan.halo && !plan.pair)) {
    fprintf(stderr, "nerfqa_b200: the frame-gradient epilogue needs the halo convolution kernel\n");
    return 0;
  }
  if (plan.pair) {
    const int rc = launch_conv3x3_tc_pair(plan, plan.sm_count, st);
    if (rc > 0) return rc;
    fprintf(stderr, "nerfqa_b200: CTA-pair conv configuration does not fit (cin %d cout %d)\n", plan.cin, plan.cout);
    return 0;
  }
  if (plan.halo) {
    const int rc = launch_conv3x3_tc_halo(plan, plan.sm_count, st);
    if (rc > 0) return rc;
    fprintf(stderr, "nerfqa_b200: halo conv configuration does not fit (cin %d cout %d)\n", plan.cin, plan.cout);
    return 0;
  }
  if (plan.epi == kEpiFrameGrad) {
    fprintf(stderr, "nerfqa_b200: the frame-gradient epilogue needs the halo convolution kernel\n");
    return 0;
  }
  if (plan.split) {
    switch (plan.block_n) {
      case 64: return tc::launch_impl<64, true>(plan, st);
      default: return tc::launch_impl<128, true>(plan, st);
    }
  }
  switch (plan.block_n) {
    case 64: return tc::launch_impl<64, false>(plan, st);
    case 128: return tc::launch_impl<128, false>(plan, st);
    default: return tc::launch_impl<256, false>(plan, st);
  }
}

}  // namespace nqa
