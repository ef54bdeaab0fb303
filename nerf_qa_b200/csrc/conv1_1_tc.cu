// Input normalisation + conv1_1 (3 -> 64 channels) + bias + ReLU on the tensor pipe.
//
// K = 27 is too thin for TMA-fed tiles, so the A operand is built in shared memory by producer warps:
// each producer thread owns one output pixel, gathers its 3x3x3 neighbourhood from the NCHW fp32 frame,
// normalises ((x - mean) / std, zero padding applied AFTER normalisation like the reference), rounds to
// bf16 and writes one 64-byte K-major row (K padded 27 -> 32) in the canonical no-swizzle UMMA layout.
// One elected thread issues two tcgen05.mma (M=128, N=64, K=16) per 128-pixel tile into a
// double-buffered TMEM accumulator; epilogue warps add the bias, apply ReLU and store NHWC bf16.
// bf16x3 mode writes hi/lo operand pairs and issues hi*hi + hi*lo + lo*hi.
//
// Reference semantics: DISTS_pt.py:92 (normalise) + vgg16.features[0:2] (DISTS_pt.py:36-37).
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace nqa {
namespace c11 {

using namespace tc;

constexpr int TILE_M = 128;
constexpr int KPAD = 32;
constexpr int NOUT = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = TILE_M * KPAD * 2;   // 8 KB
constexpr int W_BYTES = NOUT * KPAD * 2;     // 4 KB
constexpr int NUM_PRODUCER_SETS = 2;         // 2 x 128 producer threads, alternating tiles
constexpr int NUM_THREADS = (NUM_PRODUCER_SETS * 4 + 1 + 4) * 32;  // 8 producer warps, 1 MMA warp, 4 epilogue warps

__host__ __device__ inline int a_offset_bytes(int row, int kchunk) { return noswz_offset_bytes(row, kchunk); }

struct Params {
  const float* x;
  const float* y;
  int nx, n, h, w;
  int tiles_w, num_tiles;
  const bf16* wimg;   // [2][64*32] bf16 in shared-memory image order (hi, then lo)
  const float* bias;
  bf16* out;
  bf16* out_lo;
};

template <bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 2)
conv1_1_tc_kernel(const Params p) {
  constexpr int A_STAGE = A_BYTES * (SPLIT ? 2 : 1);
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET from the __shared__ array (a round trip through uintptr_t would turn every access into a generic
  // LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  uint8_t* a_smem = smem;                                   // [STAGES][A_STAGE]
  uint8_t* w_smem = smem + STAGES * A_STAGE;                // hi (+ lo)
  float* bias_s = reinterpret_cast<float*>(w_smem + 2 * W_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + NOUT);
  uint64_t* full_bar = bars;               // [STAGES]  128 producer arrivals
  uint64_t* empty_bar = bars + STAGES;     // [STAGES]  tcgen05.commit
  uint64_t* tmem_full = bars + 2 * STAGES;       // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < (SPLIT ? 2 : 1) * W_BYTES / 16; i += NUM_THREADS)
    reinterpret_cast<uint4*>(w_smem)[i] = reinterpret_cast<const uint4*>(p.wimg)[i];
  if (threadIdx.x < NOUT) bias_s[threadIdx.x] = p.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weights written by the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // =============================== im2col producers ===============================
    const int set = warp >> 2;
    const int row = threadIdx.x & 127;
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float inv_std[3] = {kInvStd0, kInvStd1, kInvStd2};
    for (int local = set;; local += NUM_PRODUCER_SETS) {
      const int tile = blockIdx.x + local * gridDim.x;
      if (tile >= p.num_tiles) break;
      const int stage = local % STAGES;
      const uint32_t phase = (local / STAGES) & 1;
      int t = tile;
      const int twi = t % p.tiles_w; t /= p.tiles_w;
      const int oy = t % p.h;
      const int img = t / p.h;
      const int ox = twi * TILE_M + row;
      const float* src = img < p.nx ? p.x + (size_t)img * 3 * p.h * p.w : p.y + (size_t)(img - p.nx) * 3 * p.h * p.w;

      // All 27 loads are issued unconditionally from clamped addresses and normalised / zeroed afterwards: a load
      // under a branch makes the compiler wait for it at the reconvergence point (27 serial latencies per pixel).
      float v[KPAD];
      uint32_t okmask = 0;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int iy = oy - 1 + r;
          const bool rowok = iy >= 0 && iy < p.h;
          const float* rp = src + ((size_t)ci * p.h + min(max(iy, 0), p.h - 1)) * p.w;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int ix = ox - 1 + s;
            if (rowok && ix >= 0 && ix < p.w) okmask |= 1u << ((ci * 3 + r) * 3 + s);
            v[(ci * 3 + r) * 3 + s] = __ldg(rp + min(max(ix, 0), p.w - 1));
          }
        }
      }
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int k = ci * 9; k < ci * 9 + 9; ++k) v[k] = ((okmask >> k) & 1) ? (v[k] - mean[ci]) * inv_std[ci] : 0.f;
#pragma unroll
      for (int k = 27; k < KPAD; ++k) v[k] = 0.f;

      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* a_hi = a_smem + stage * A_STAGE;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint4 hi;
        hi.x = Vec<bf16>::pack2(v[kc * 8 + 0], v[kc * 8 + 1]);
        hi.y = Vec<bf16>::pack2(v[kc * 8 + 2], v[kc * 8 + 3]);
        hi.z = Vec<bf16>::pack2(v[kc * 8 + 4], v[kc * 8 + 5]);
        hi.w = Vec<bf16>::pack2(v[kc * 8 + 6], v[kc * 8 + 7]);
        *reinterpret_cast<uint4*>(a_hi + a_offset_bytes(row, kc)) = hi;
        if (SPLIT) {
          float l[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) l[j] = v[kc * 8 + j] - __bfloat162float(__float2bfloat16_rn(v[kc * 8 + j]));
          uint4 lo;
          lo.x = Vec<bf16>::pack2(l[0], l[1]);
          lo.y = Vec<bf16>::pack2(l[2], l[3]);
          lo.z = Vec<bf16>::pack2(l[4], l[5]);
          lo.w = Vec<bf16>::pack2(l[6], l[7]);
          *reinterpret_cast<uint4*>(a_hi + A_BYTES + a_offset_bytes(row, kc)) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // make the rows visible to the MMA (async proxy)
      mbar_arrive(&full_bar[stage]);
    }
  } else if (warp == 8) {
    // =============================== MMA issuer (whole warp, one elected lane issues) ===============================
    constexpr uint32_t idesc = make_idesc(TILE_M, NOUT);
    const uint32_t w_hi = smem_u32(w_smem), w_lo = w_hi + W_BYTES;
    for (int local = 0;; ++local) {
      const int tile = blockIdx.x + local * gridDim.x;
      if (tile >= p.num_tiles) break;
      const int stage = local % STAGES;
      const uint32_t phase = (local / STAGES) & 1;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(a_smem + stage * A_STAGE), a_lo = a_hi + A_BYTES;
        const uint32_t d_tmem = tmem_base + acc * NOUT;
#pragma unroll
        for (int k = 0; k < KPAD / 16; ++k) {
          const uint32_t koff = k * 256;   // two 8-element core matrices per UMMA_K
          umma_bf16(d_tmem, make_desc_noswz(a_hi + koff), make_desc_noswz(w_hi + koff), idesc, k > 0 ? 1u : 0u);
          if (SPLIT) {
            umma_bf16(d_tmem, make_desc_noswz(a_hi + koff), make_desc_noswz(w_lo + koff), idesc, 1u);
            umma_bf16(d_tmem, make_desc_noswz(a_lo + koff), make_desc_noswz(w_hi + koff), idesc, 1u);
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue ===============================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    for (int local = 0;; ++local) {
      const int tile = blockIdx.x + local * gridDim.x;
      if (tile >= p.num_tiles) break;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      int t = tile;
      const int twi = t % p.tiles_w; t /= p.tiles_w;
      const int oy = t % p.h;
      const int img = t / p.h;
      const int ox = twi * TILE_M + row;
      const bool valid = ox < p.w;
      const size_t obase = (((size_t)img * p.h + oy) * p.w + ox) * NOUT;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < NOUT / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * NOUT + ch * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r[j]) + bias_s[ch * 32 + j], 0.f);
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + obase + ch * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = Vec<bf16>::pack2(v[q * 8 + 0], v[q * 8 + 1]);
            o.y = Vec<bf16>::pack2(v[q * 8 + 2], v[q * 8 + 3]);
            o.z = Vec<bf16>::pack2(v[q * 8 + 4], v[q * 8 + 5]);
            o.w = Vec<bf16>::pack2(v[q * 8 + 6], v[q * 8 + 7]);
            dst[q] = o;
          }
          if (SPLIT) {
            uint4* dlo = reinterpret_cast<uint4*>(p.out_lo + obase + ch * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float l[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) l[j] = v[q * 8 + j] - __bfloat162float(__float2bfloat16_rn(v[q * 8 + j]));
              uint4 o;
              o.x = Vec<bf16>::pack2(l[0], l[1]);
              o.y = Vec<bf16>::pack2(l[2], l[3]);
              o.z = Vec<bf16>::pack2(l[4], l[5]);
              o.w = Vec<bf16>::pack2(l[6], l[7]);
              dlo[q] = o;
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
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// weight image: element (co, k) of the [64 x 32] K-major operand, hi and lo halves
__global__ void pack_w11_tc_kernel(const float* __restrict__ w /*[64][27] OIHW*/, bf16* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NOUT * KPAD) return;
  const int co = i / KPAD, k = i % KPAD;
  const float v = k < 27 ? w[co * 27 + k] : 0.f;
  const bf16 hi = __float2bfloat16_rn(v);
  const int off = (a_offset_bytes(co, k >> 3) >> 1) + (k & 7);
  img[off] = hi;
  img[NOUT * KPAD + off] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

}  // namespace c11

int launch_pack_w11_tc(const float* w_oihw, bf16* img, cudaStream_t st) {
  c11::pack_w11_tc_kernel<<<(c11::NOUT * c11::KPAD + 255) / 256, 256, 0, st>>>(w_oihw, img);
  return 1;
}

int launch_conv1_1_tc(const float* x, const float* y, int nx, int n, int h, int w, const bf16* wimg, const float* bias,
                      bf16* out, bf16* out_lo, int sm_count, cudaStream_t st) {
  using namespace c11;
  Params p;
  p.x = x; p.y = y; p.nx = nx; p.n = n; p.h = h; p.w = w;
  p.tiles_w = ceil_div(w, TILE_M);
  p.num_tiles = n * h * p.tiles_w;
  p.wimg = wimg; p.bias = bias; p.out = out; p.out_lo = out_lo;
  const bool split = out_lo != nullptr;
  const int a_stage = A_BYTES * (split ? 2 : 1);
  const int smem = STAGES * a_stage + 2 * W_BYTES + NOUT * 4 + 256 + 128;
  const int grid = p.num_tiles < 2 * sm_count ? p.num_tiles : 2 * sm_count;   // two resident CTAs per SM
  if (split) {
    ensure_dynamic_smem(conv1_1_tc_kernel<true>, smem);
    conv1_1_tc_kernel<true><<<grid, NUM_THREADS, smem, st>>>(p);
  } else {
    ensure_dynamic_smem(conv1_1_tc_kernel<false>, smem);
    conv1_1_tc_kernel<false><<<grid, NUM_THREADS, smem, st>>>(p);
  }
  return 1;
}

}  // namespace nqa
