// Host-side launchers of the nerfqa_b200 kernels.  Every launcher enqueues on `st`, returns the number
// of kernels it launched, and never synchronises.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace nqa {

typedef __nv_bfloat16 bf16;

// ---- statistics -----------------------------------------------------------------------------------
// Partial sums of one pyramid level: partials[((b*chunks + chunk)*5 + s)*C + c], s in
// {sum x', sum y', sum x'^2, sum y'^2, sum x'*y'}; fp32 per chunk, combined in fp64 by the finalize kernel.
// x' = x - pivot_x, y' = y - pivot_y with the channel's FIRST pixel as pivot (stored behind the partial sums at
// partials + pairs*chunks*5*C + (b*2 + {0,1})*C + c): variances and covariances are shift invariant, the shifted
// one-pass formula has no cancellation against the mean (a constant map gives exactly zero like the reference's
// two-pass variance, DISTS_pt.py:137-139), and the mean is pivot + E[x'].
struct LevelPartials {
  const float* partials;
  const float* piv;   // [pairs][2][C] pivots (first pixel of x / y per channel): the sums are of (x - piv_x), (y - piv_y)
  int chunks;
  int C;
  int choff;     // first channel of this level inside the 1475-vector
  double inv_n;  // 1 / (H_k * W_k)
};
struct FinalizeParams {
  LevelPartials lv[6];
};

int stats_chunks_nhwc(int hw, int c, int elem_bytes, int pairs);
int stats_chunks_planar(int hw, int c, int pairs);

// fx_lo/fy_lo != nullptr: bf16x3 carrier, value = hi + lo
template <typename T>
int launch_stats_nhwc(const T* fx, const T* fy, const T* fx_lo, const T* fy_lo, int pairs, int hw, int c, int chunks,
                      float* partials, cudaStream_t st);
// NCHW fp32 planes (level 0 and the from_feats API): x[b][c][hw]
int launch_stats_planar(const float* x, const float* y, int pairs, int c, int hw, int chunks, float* partials,
                        cudaStream_t st);
// score[b] = 1 - sum_c (alpha_c/w S1 + beta_c/w S2); optional s1s2[b][2][1475]; optional moments[b][1475][5]
// (mean x, mean y, var x, var y, cov) for the backward pass.
// terms: [pairs][1475] fp64 scratch
int launch_dists_finalize(const FinalizeParams& p, const float* alpha, const float* beta, bool normalize, int pairs,
                          float* score, float* s1s2, float* moments, double* terms, cudaStream_t st);

// ---- pooling / layout -----------------------------------------------------------------------------
template <typename T>
int launch_l2pool_nhwc(const T* in, const T* in_lo, T* out, T* out_lo, int n, int h, int w, int c, cudaStream_t st);
// L2pooling of x and y features fused with the level's DISTS moment sums (one read of the level).
// chunks = pool_stats_chunks(h, w, c, sizeof(T)) partial-sum slots per pair.
int pool_stats_chunks(int h, int w, int c, int elem_bytes);
void pool_stats_plan(int h, int w, int c, int elem_bytes, int* rows_per_block, int* col_blocks, int* strips);
// TMA-fed version for bf16 levels with x and y in one tensor (fy == fx + pairs*h*w*c): same results bit for bit
// (pool_stats_tma.cu).  Returns 0 if the configuration is not served.
int launch_pool_stats_tma(const bf16* fx, const bf16* fy, bf16* px_out, bf16* py_out, int pairs, int h, int w, int c, int chunks,
                          float* partials, cudaStream_t st);
template <typename T>
int launch_pool_stats(const T* fx, const T* fy, T* px_out, T* py_out, int pairs, int h, int w, int c, int chunks,
                      float* partials, cudaStream_t st, bool allow_tma = true);
int launch_l2pool_nchw(const float* in, float* out, int n, int c, int h, int w, cudaStream_t st);
template <typename T>
int launch_nhwc_to_nchw(const T* in, float* out, int n, int c, int h, int w, cudaStream_t st);
template <typename T>
int launch_nchw_to_nhwc(const float* in, T* out, T* out_lo, int n, int c, int h, int w, cudaStream_t st);
int launch_nhwc_to_nchw_bf16(const bf16* in, bf16* out, int n, int c, int h, int w, cudaStream_t st);
// bf16x3 carrier: hi = bf16(v), lo = bf16(v - hi)
int launch_nhwc_split_to_nchw(const bf16* hi, const bf16* lo, float* out, int n, int c, int h, int w, cudaStream_t st);

// ---- convolutions ---------------------------------------------------------------------------------
// Epilogue of the 3x3 convolution kernels.  Forward: bias + ReLU.  Backward (data gradient = the same convolution
// with rotated / transposed weights): either the plain sum (gradient w.r.t. a pooled tensor) or the sum masked by
// the ReLU of the layer below (mask = that layer's saved output, same shape as this kernel's output).
// kEpiDivide: the sum divided by a saved tensor of the output's shape -- the data gradient w.r.t. an L2-pooled tensor
// is only ever used as grad / pooled (backward of the square root), so the quotient is formed once here instead of in
// each of the up to nine input pixels of the pooling window.
// kEpiFrameGrad (halo kernel, 64-channel tile): backward of conv1_1 + input normalisation.  The weights are the
// rotated conv1_1 filters with the three input channels as output rows 0..2 (rows 3..63 zero); the epilogue writes
// sum / std[ci] as fp32 PLANES into the frame gradients (images [0,nx) -> fg_x, the rest -> fg_y) instead of an NHWC
// tensor.
enum ConvEpilogue { kEpiBiasRelu = 0, kEpiPlain = 1, kEpiMask = 2, kEpiDivide = 3, kEpiFrameGrad = 4 };
// conv1_1 fused with the input normalisation: images [0,nx) come from x, [nx,n) from y (NCHW fp32);
// weights packed [27][64] (k = (ci*3 + r)*3 + s), output NHWC [n,h,w,64].  If out_lo != nullptr the output is
// written as a bf16 hi/lo pair (T must be bf16).
template <typename T>
int launch_conv1_1(const float* x, const float* y, int nx, int n, int h, int w, const float* wpack, const float* bias,
                   T* out, T* out_lo, cudaStream_t st);
// tensor-pipe version (bf16 / bf16x3): im2col rows built by producer warps, K = 27 padded to 32.
// wimg: [2][64*32] bf16 shared-memory image of the weights (hi, lo) from launch_pack_w11_tc.
int launch_pack_w11_tc(const float* w_oihw, bf16* img, cudaStream_t st);
int launch_conv1_1_tc(const float* x, const float* y, int nx, int n, int h, int w, const bf16* wimg, const float* bias,
                      bf16* out, bf16* out_lo, int sm_count, cudaStream_t st);
// strict fp32 CUDA-core conv: NHWC fp32, weights [9][cin][cout]
int launch_conv3x3_f32(const float* in, float* out, const float* wpack, const float* bias, int n, int h, int w, int cin,
                       int cout, cudaStream_t st, int epi = kEpiBiasRelu, const float* mask = nullptr);

// tcgen05 implicit-GEMM conv: NHWC bf16 activations, weights [cout][9][cin] bf16 (K-major), fp32 bias.
// split = bf16x3 mode (hi/lo pairs for activations and weights, three MMAs per k-step).
// Kernel-selection switches of the tensor-pipe convolutions, read from the environment ONCE per context
// (nqa_ctx_create) and carried by the context -- nothing about the selection is process-global.
struct ConvTcConfig {
  int sm_count = 148;
  int halo_mode = 1;        // NQA_CONV_HALO: 0 = per-tap kernel everywhere; 1..4 = halo kernel with (halo width, base mode)
  int halo_max_cin = 512;   // NQA_CONV_HALO_MAXCIN
  bool n64 = true;          // NQA_CONV_N64: 128 -> 128 channels as two resident 64-channel tiles (halo kernel)
  bool pair = true;         // NQA_CONV_PAIR: CTA pairs (tcgen05 cta_group::2) for the wide layers
  int pair_min_cin = 128;   // NQA_PAIR_MINCIN
  bool split_pair = true;   // NQA_PAIR_SPLIT: bf16x3 forward convolutions on the CTA-pair kernel (hi + lo operands)
};
ConvTcConfig conv_tc_config_from_env(int sm_count);

struct ConvTcPlan {
  CUtensorMap tmap_a, tmap_a_lo, tmap_b, tmap_b_lo;
  CUtensorMap tmap_out;   // halo kernel: output [n,h,w,cout] with box {64, 8, 16, 1} for the TMA-store epilogue
  CUtensorMap tmap_out_lo;  // bf16x3 on the CTA-pair kernel: the residual tensor
  int n, h, w, cin, cout;
  int sm_count;        // persistent-grid size (the context's device)
  int tw_log2;
  int tiles_w, tiles_h;
  int block_n;
  bool split;
  bool pair;           // CTA-pair kernel (conv_tc_pair.cu): halo reuse + tcgen05 cta_group::2, weights box = block_n/2 rows
  bool halo;           // halo-reuse kernel (conv_tc_halo.cu): tile 16x8, A halo loaded once per 64-channel chunk
  int halo_hbw;        // halo box width in pixels
  int halo_desc_mode;  // how the UMMA descriptor's base-offset field is filled
  const float* bias;
  bf16* out;
  bf16* out_lo;
  bool strips_out;     // pair kernel: `out` is a strip-plane buffer [n][cout][w/8][h][8] (tmap_out = make_strip_out_map)
  int epi;             // ConvEpilogue (halo / pair kernels; the plain kernel is forward-only)
  const bf16* mask;    // kEpiMask / kEpiDivide: saved activation [n,h,w,cout]
  const bf16* mask_lo; // bf16x3 carrier of the saved activation (per-tap kernel, kEpiDivide)
  float* fg_x;         // kEpiFrameGrad: frame gradients, NCHW fp32 [nx,3,h,w] / [n-nx,3,h,w]
  float* fg_y;
  int fg_nx;
};
// Fills the tensor maps; returns 0 on success, nonzero + message on failure.
int conv_tc_make_plan(ConvTcPlan* plan, const ConvTcConfig& cfg, const bf16* in, const bf16* in_lo, const bf16* wk,
                      const bf16* wk_lo, const float* bias, bf16* out, bf16* out_lo, int n, int h, int w, int cin, int cout,
                      bool split, const char** err, bool strips_out = false, bool split_on_pairs = false);
// split_on_pairs: a bf16x3 FORWARD convolution may use the CTA-pair kernel (its data-gradient epilogues exist on the per-tap
// kernel only)
// whether conv_tc_make_plan picks the CTA-pair kernel for a layer (the kernel that can write strip planes)
bool conv_tc_uses_pair(const ConvTcConfig& cfg, int cin, int cout, bool split);
int launch_conv3x3_tc(const ConvTcPlan& plan, cudaStream_t st);
int launch_conv3x3_tc_halo(const ConvTcPlan& plan, int sm_count, cudaStream_t st);
int launch_conv3x3_tc_pair(const ConvTcPlan& plan, int sm_count, cudaStream_t st);
// 4-D SWIZZLE_128B tensor map over NHWC bf16 activations with box {64 channels, box_w, box_h, 1}; 0 on success.
int conv_tc_make_act_map(CUtensorMap* map, const bf16* ptr, int n, int h, int w, int c, int box_w, int box_h);
// 2-D SWIZZLE_128B tensor map over K-major weights [cout][k] bf16 with box {64, block_n}; 0 on success.
int conv_tc_make_weight_map(CUtensorMap* map, const bf16* wk, int cout, int k, int block_n);
// 3-D SWIZZLE_64B tensor map over bf16 planes [planes][h][w] (w % 8 == 0) with box {32 columns, box_h rows, 1}; out-of-image
// parts of a box are zero-filled.  0 on success.
int make_plane_map_sw64(CUtensorMap* map, const bf16* ptr, int planes, int h, int w, int box_h);
// Strip layout of bf16 planes, [plane][w / 8][h][8] (w % 8 == 0): 8-pixel-wide column strips with the rows of a strip
// contiguous, so that a 16 x 8 pixel tile of one channel is 256 contiguous bytes (what a convolution epilogue holds per
// channel) and 8 rows of a strip are 128 contiguous bytes (what an ldmatrix of the window kernel reads).  3-D tensor map
// {8 h, w / 8, planes} with box {box_elems, box_strips, 1}, no swizzle; 0 on success.
int make_strip_map(CUtensorMap* map, const bf16* ptr, int planes, int h, int w, int box_elems, int box_strips);
// The same layout seen from a convolution epilogue: 4-D map {8 h, w / 8, c, n} with box {128, 1, 64, 1} -- the 16 x 8
// pixel tile of 64 channels, staged as [channel][pixel], leaves as 64 runs of 256 contiguous bytes; rows beyond the
// image are clipped.  0 on success.
int make_strip_out_map(CUtensorMap* map, const bf16* ptr, int n, int c, int h, int w);

// Fused VGG stage 1 (bf16 trunk): normalise + conv1_1 + ReLU + conv1_2 + ReLU, relu1_1 stays on chip
// (stage1_fused.cu).  w1img from launch_pack_w11_tc, w2map from conv_tc_make_weight_map(.., 64, 576, 64).
int launch_stage1_fused(const float* x, const float* y, int nx, int n, int h, int w, const bf16* w1img,
                        const float* bias1, const CUtensorMap& w2map, const float* bias2, bf16* out, int sm_count, bool pair,
                        bool a2_linear, cudaStream_t st, bool strips_out = false);   // strips_out: `out` = strip planes

// ---- backward (backward.cu) -------------------------------------------------------------------------
struct GradLevels {
  double inv_n[6];   // 1 / (H_k * W_k)
};
// Coefficient table coef[b][q][kCoefStride], q in (Px, Qx, R, Py, Qy): d score_b/d x_i = Px + Qx x_i + R y_i (times
// grad_score[b], already folded in).  Level k's channels start at coef_offset(k), padded to multiples of 8 so that the
// NHWC kernel reads its 4 / 8 channels as 16-byte vectors.  Channels of levels >= levels_valid get zeros.
constexpr int kCoefStride = 1536;
__host__ __device__ constexpr int coef_offset(int k) { return k == 0 ? 0 : k == 1 ? 8 : k == 2 ? 72 : k == 3 ? 200 : k == 4 ? 456 : 968; }
int launch_stat_coef(const float* moments, const float* alpha, const float* beta, bool normalize, const float* grad_score,
                     const GradLevels& lv, int levels_valid, int pairs, float* coef, cudaStream_t st);
// NCHW fp32 planes of one level: gx (+)= Px + Qx x + R y, gy (+)= Py + Qy y + R x
int launch_stat_grad_planar(const float* x, const float* y, const float* coef, int level, int pairs, int c, int hw,
                            float* gx, float* gy, bool accumulate, cudaStream_t st);
// NHWC level: statistics gradient (+ L2pooling backward when qpoolx != nullptr), masked by the level's ReLU.
// qpoolx / qpooly: (gradient w.r.t. the pooled tensor) / (pooled tensor), as written by a kEpiDivide convolution.
// lo (bf16x3 carriers, T = bf16): residual tensors of fx / fy / qpoolx / qpooly and of the two outputs, all or none.
template <typename T>
struct LevelGradLo {
  const T* fx = nullptr; const T* fy = nullptr; const T* qpoolx = nullptr; const T* qpooly = nullptr;
  T* gx = nullptr; T* gy = nullptr;
};
template <typename T>
int launch_level_grad(const T* fx, const T* fy, const T* qpoolx, const T* qpooly, const float* coef, int level, int pairs,
                      int h, int w, int c, T* gx, T* gy, cudaStream_t st, const LevelGradLo<T>* lo = nullptr);
// conv1_1 + normalisation backward: gz [n,h,w,64] -> NCHW fp32 frame gradients.  The 9x64x3 weights travel as a kernel
// parameter (constant bank: every thread reads the same weight at the same time, no shared-memory traffic).
struct Conv11BwdWeights { float w[9][64][3]; };
template <typename T>
int launch_conv1_1_bwd(const T* gz, const Conv11BwdWeights& wb, int nx, int n, int h, int w, float* gx, float* gy,
                       cudaStream_t st, const T* gz_lo = nullptr);
int launch_pack_w_bwd(const float* w_oihw, int cin, int cout, float* w_f32, bf16* w_hi, cudaStream_t st, bf16* w_lo = nullptr);
int launch_pack_w11_bwd(const float* w_oihw, float* out, cudaStream_t st);

// ---- frame resizing (resize.cu): NCHW fp32 planes [planes][h][w] -> [planes][oh][ow], bilinear, align_corners=False
int launch_resize_bilinear(const float* in, float* out, int planes, int h, int w, int oh, int ow, bool antialias,
                           cudaStream_t st);

// ---- micro-benchmarks (microbench.cu): measured roofline denominators ------------------------------------------
// which = 0: packed fp32 FMAs (FFMA2) on the CUDA cores; which = 1: warp-level mma.sync m16n8k16 bf16.  FMA/s.
int microbench_rate(int which, int sm_count, float* scratch, cudaStream_t st, double* fma_per_s);

// ---- A-DISTS --------------------------------------------------------------------------------------
// See adists.cu.
struct AdistsLevel {
  int c, h, w;       // level shape
  int oh, ow;        // window-map shape (h-win+1, w-win+1) or 1x1 in the global-statistics fallback
  bool windowed;
  int choff;
};

}  // namespace nqa
