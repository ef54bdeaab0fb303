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
// fx[k], fy[k]: [pairs, C_k, H_k, W_k] fp32 (level 0 = raw frames).  score: [pairs].  Returns the number
// of kernels launched.  map_out (optional): [pairs, H, W] full-resolution map 1 - sum_k bilinear(D_map_k) (as_map=True).
int adists_stage(const float* const fx[6], const float* const fy[6], int pairs, int H, int W, int window, void* ws,
                 float* score, cudaStream_t st, float* map_out = nullptr);

}  // namespace nqa
