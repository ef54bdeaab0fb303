// Tensor-pipe (mma.sync, banded Toeplitz) version of the A-DISTS window statistics for bf16 feature planes; see
// adists_tc.cu.  Same outputs as window_kernel in adists.cu: per frame pair and channel split the three maps
// sum_c w_c T_c, sum_c w_c S_c and sum_c gamma_c of shape (h - window + 1) x (w - window + 1).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace nqa {

// bf16 planes with 16-byte aligned rows (w % 8 == 0) and an odd window <= 21
bool window_mma_supported(int window, int h, int w, const void* fx, const void* fy);
// block tiling of the output map of an h x w level (64 x 72 or 64 x 48 outputs per block)
void window_mma_tiles(int h, int w, int window, int* tiles_x, int* tiles_y);
// taps_f32: the `window` fp32 gaussian taps (rounded to bf16 inside; the result is renormalised by their exact sum)
// strips: the planes are in the strip layout [plane][w / 8][h][8] (kernels.h make_strip_map).  Returns 1, 0 when no tensor
// map could be made for planar planes (the caller uses the CUDA-core kernel), -1 on failure with strip planes.
int launch_window_mma(const __nv_bfloat16* fx, const __nv_bfloat16* fy, int pairs, int c, int h, int w, int window, int csplit,
                      int choff, const double* chan, const float* wt, const float* taps_f32, long long one_term_min_pixels, bool strips,
                      float* amap, float* bmap, float* gmap, cudaStream_t st);   // one_term_min_pixels < 0: always two terms

}  // namespace nqa
