// A-DISTS statistics stage: everything of ADISTS.forward after the trunk
// (nerf_qa/ADISTS/ADISTS.py:147-197) on planar (NCHW) fp32 feature levels.
#pragma once

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
int adists_stage(const void* const fx[6], const void* const fy[6], bool feats_bf16, int pairs, int H, int W, int window,
                 void* ws, float* score, cudaStream_t st, float* map_out = nullptr);

}  // namespace nqa
