/*
 * A plain-C consumer of the nerfqa_b200 C-ABI: what a non-Python caller of the hot path links against.
 * Scores B synthetic frame pairs from HOST buffers with nqa_dists_forward_host (the call that replaces
 * `dists_model(ref.to(device), render.to(device)).cpu()` of prep.py:186-190 in the reference) and checks the two
 * size-independent properties of the metric: DISTS(x, x) == 0 and symmetry.
 *
 *   gcc -std=c99 -Iinclude examples/score_frames.c -Lnerf_qa_b200 -lnerfqa_b200 -Wl,-rpath,$PWD/nerf_qa_b200 -lm -o score_frames
 *   ./score_frames [B H W]
 *
 * Exit code 0 = scores printed and properties hold, 2 = no CUDA device (nothing computed: there is no CPU fallback),
 * 1 = any other failure.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "nerfqa_b200.h"

static const int kCin[NQA_NUM_CONVS] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512};
static const int kCout[NQA_NUM_CONVS] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};

static unsigned int g_seed = 12345u;
static float frand(void) { /* uniform in [0, 1) */
  g_seed = g_seed * 1664525u + 1013904223u;
  return (float)(g_seed >> 8) * (1.0f / 16777216.0f);
}

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != NQA_OK) {                                                             \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, nqa_last_error());         \
      return rc_ == NQA_ERR_CUDA ? 2 : 1;                                            \
    }                                                                                \
  } while (0)

int main(int argc, char** argv) {
  const int B = argc > 3 ? atoi(argv[1]) : 2, H = argc > 3 ? atoi(argv[2]) : 96, W = argc > 3 ? atoi(argv[3]) : 128;
  printf("nerfqa_b200 version %d\n", nqa_version());
  nqa_ctx* ctx = NULL;
  CHECK(nqa_ctx_create(0, &ctx));

  /* He-initialised VGG16 trunk (the reference downloads ImageNet weights; any weights exercise the same path) */
  float* w[NQA_NUM_CONVS];
  float* b[NQA_NUM_CONVS];
  const float* wc[NQA_NUM_CONVS];
  const float* bc[NQA_NUM_CONVS];
  for (int i = 0; i < NQA_NUM_CONVS; ++i) {
    const size_t nw = (size_t)kCout[i] * kCin[i] * 9;
    const float scale = sqrtf(6.0f / (float)(kCin[i] * 9));
    w[i] = (float*)malloc(nw * sizeof(float));
    b[i] = (float*)malloc((size_t)kCout[i] * sizeof(float));
    if (!w[i] || !b[i]) return 1;
    for (size_t j = 0; j < nw; ++j) w[i][j] = (2.0f * frand() - 1.0f) * scale;
    for (int j = 0; j < kCout[i]; ++j) b[i][j] = 0.05f * frand();
    wc[i] = w[i];
    bc[i] = b[i];
  }
  CHECK(nqa_set_vgg_weights(ctx, wc, bc));
  float* alpha = (float*)malloc(NQA_NUM_CHANNELS * sizeof(float));
  float* beta = (float*)malloc(NQA_NUM_CHANNELS * sizeof(float));
  for (int c = 0; c < NQA_NUM_CHANNELS; ++c) {
    alpha[c] = 0.05f + 0.1f * frand();
    beta[c] = 0.05f + 0.1f * frand();
  }
  CHECK(nqa_set_alpha_beta(ctx, alpha, beta));

  const size_t n = (size_t)B * 3 * H * W;
  float* x = (float*)malloc(n * sizeof(float));
  float* y = (float*)malloc(n * sizeof(float));
  float* sxy = (float*)malloc((size_t)B * sizeof(float));
  float* syx = (float*)malloc((size_t)B * sizeof(float));
  float* sxx = (float*)malloc((size_t)B * sizeof(float));
  if (!x || !y || !sxy || !syx || !sxx) return 1;
  for (size_t i = 0; i < n; ++i) {
    x[i] = frand();
    y[i] = fminf(fmaxf(x[i] + 0.2f * (frand() - 0.5f), 0.0f), 1.0f);
  }
  CHECK(nqa_dists_forward_host(ctx, x, y, B, H, W, NQA_PREC_BF16, sxy));
  CHECK(nqa_dists_forward_host(ctx, y, x, B, H, W, NQA_PREC_BF16, syx));
  CHECK(nqa_dists_forward_host(ctx, x, x, B, H, W, NQA_PREC_BF16, sxx));
  int ok = 1;
  for (int i = 0; i < B; ++i) {
    printf("pair %d: DISTS(x,y) = %.6f  DISTS(y,x) = %.6f  DISTS(x,x) = %.3e\n", i, sxy[i], syx[i], sxx[i]);
    if (!(sxy[i] > 0.0f && sxy[i] < 1.0f) || fabsf(sxy[i] - syx[i]) > 1e-5f || fabsf(sxx[i]) > 1e-5f) ok = 0;
  }
  CHECK(nqa_ctx_destroy(ctx));
  for (int i = 0; i < NQA_NUM_CONVS; ++i) { free(w[i]); free(b[i]); }
  free(alpha); free(beta); free(x); free(y); free(sxy); free(syx); free(sxx);
  printf(ok ? "OK\n" : "PROPERTY CHECK FAILED\n");
  return ok ? 0 : 1;
}
