// A-DISTS statistics stage: everything of ADISTS.forward after the trunk
// (nerf_qa/ADISTS/ADISTS.py:147-197) on planar (NCHW) fp32 feature levels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace nqa {

constexpr int kAdistsMaxWindow = 33;

// true if this build has a window kernel for `window`
bool adists_window_supported(int window);
size_t adists_workspace_bytes(int pairs, int H, int W, int window);
// fx[k], fy[k]: planes [pairs, C_k, H_k, W_k]; level 0 (the raw frames) is fp32, levels 1..5 are fp32 or -- when
// feats_bf16 -- bf16 (the bf16 trunk's activations, half the bytes).  score: [pairs].  Returns the number of kernels
// launched.  map_out (optional): [pairs, H, W] full-resolution map 1 - sum_k bilinear(D_map_k) (as_map=True).
// Optional copies of the stage's intermediates (device pointers, any may be null): rows a6 / a7 / a8 of the scope
// table are checked level by level against the reference with these.
struct AdistsParts {
  float* weights = nullptr;       // [pairs][1475] clamped entropy channel weights (ADISTS.py:152-160)
  float* level_terms = nullptr;   // [pairs][6] D_k = mean_hw sum_c w_c((1-ps) T + ps S) (ADISTS.py:186-191)
  float* ps[6] = {};              // ps_prod maps [pairs][oh_k*ow_k] (compute_prob, ADISTS.py:71-100)
};
// Optional timing hook of the caller (per-family CUDA events): begin(user, part) / end(user, launches) bracket the launches of
// part 0 = reductions, entropy weights and the ps chain, part 1 = the window kernels.
// Kernel selection of the stage, owned by the context (read from the environment once, at nqa_ctx_create).
struct AdistsOptions {
  bool use_mma = true;   // NQA_ADISTS_MMA=0: CUDA-core window kernel for bf16 planes too
  int mma_terms = 0;     // NQA_ADISTS_MMA_TERMS: 0 = one bf16 term per operand for maps >= one_term_min_pixels, else two; 2 = always two
  long long one_term_min_pixels = 100000;   // NQA_ADISTS_ONE_TERM_MIN_PIXELS (see adists_tc.cu launch_window_mma)
};
struct AdistsProf {
  void* user = nullptr;
  void (*begin)(void* user, int part) = nullptr;
  void (*end)(void* user, int launches) = nullptr;   // launches made since begin
};
int adists_stage(const void* const fx[6], const void* const fy[6], bool feats_bf16, int pairs, int H, int W, int window,
                 void* ws, float* score, cudaStream_t st, float* map_out = nullptr, const AdistsParts* parts = nullptr,
                 const AdistsProf* prof = nullptr, unsigned sums_done_mask = 0, const AdistsOptions* opt = nullptr,
                 unsigned strips_mask = 0);   // strips_mask: levels whose bf16 planes are in the strip layout
// bf16 trunk: level `level`'s NHWC activations [2*pairs][h][w][c] (x images first) -> bf16 planes [2*pairs][c][h*w],
// with the level's per-channel sums written into the workspace `ws` on the way (one read of the level instead of
// three).  Pass bit `level` in adists_stage's sums_done_mask afterwards.  Levels 1..5 (c % 64 == 0) with at least 256 pixels.
bool adists_planes_sums_supported(int level, int H, int W);
int adists_planes_and_sums(const __nv_bfloat16* nhwc, __nv_bfloat16* planes, int pairs, int level, int H, int W, int window,
                           void* ws, cudaStream_t st, bool strips = false);   // strips: write the strip layout (validation path)
// Strip planes (bf16 trunk): level `level`'s features as written by the convolution epilogues, [2*pairs][c][w/8][h][8]
// (kernels.h make_strip_map).  adists_pool_strips_and_sums reads them once: L2-pooled NHWC input of the next stage
// (pooled_out, may be null for level 5) + the level's channel sums into the workspace; pass bit `level` in BOTH masks of
// adists_stage afterwards.  Supported for levels whose width is a multiple of 8 and that are windowed (the tensor-pipe
// window kernel is the only reader of strip planes).
bool adists_strips_supported(int level, int H, int W, int window);
// dists_partials (optional): the level's pivot-shifted DISTS moments as well, [pairs][adists_strips_chunks][5][c] floats +
// [pairs][2][c] pivots (the layout dists_finalize reads), for calls that score both metrics from one trunk pass.
int adists_pool_strips_and_sums(const __nv_bfloat16* planes, __nv_bfloat16* pooled_out, int pairs, int level, int H, int W,
                                int window, void* ws, cudaStream_t st, float* dists_partials = nullptr);
int adists_strips_chunks(int c, int h, int w);
void adists_strips_plan(int c, int h, int w, int* spb, int* n_sg, int* chunks);
int launch_pool_strips_sums(const __nv_bfloat16* planes, __nv_bfloat16* pooled_out, int pairs, int c, int h, int w, float* partials,
                            float* dists_partials, cudaStream_t st);
// number of elements of the level-k ps map of one frame (1 for a fallback level)
size_t adists_ps_elems(int H, int W, int window, int level);

}  // namespace nqa
