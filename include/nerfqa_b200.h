/*
 * nerfqa_b200 — C-ABI of the B200-native DISTS / A-DISTS frame-pair scorer.
 *
 * This is the drop-in boundary for the full-reference perceptual-distance hot path of
 * kobejean/nerf-qa.  Every entry point cites the reference interface it replaces (paths relative to
 * the reference checkout).  The reference is pure Python over torch, so what a maintainer binds here
 * is exactly what nerf_qa/DISTS_pytorch/DISTS_pt.py and nerf_qa/ADISTS/ADISTS.py do with ATen today;
 * INTEGRATION.md shows the ctypes / torch glue.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / CUDA types in signatures (`stream` is a
 *     cudaStream_t passed as void*; NULL = the legacy default stream);
 *   - every call returns 0 (NQA_OK) or a negative nqa_status; nqa_last_error() gives the message
 *     of the calling thread's last failure; no exception crosses the boundary;
 *   - image tensors are NCHW fp32, values nominally in [0,1], x and y of identical shape
 *     (ADISTS.py:138 asserts it; DISTS_pt.py relies on it);
 *   - "device" pointers must live on the context's GPU; "any" pointers may be host or device (UVA);
 *   - the caller owns all input/output buffers and the stream; the context owns packed weights and
 *     a workspace arena that grows on demand (no allocation after warm-up at a fixed shape);
 *   - calls on one context are not thread-safe (same contract as a torch nn.Module); use one
 *     context per GPU, one process (or thread) per GPU;
 *   - calls on one context may use different streams: the context orders them itself (an event is recorded on
 *     the stream of every call and the next call on another stream waits for it), because the workspace arena, the
 *     channel weights and the staging buffers are shared; independent work belongs on separate contexts;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with NQA_ERR_CUDA.
 */
#ifndef NERFQA_B200_H
#define NERFQA_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define NQA_API __attribute__((visibility("default")))
#else
#define NQA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define NQA_NUM_CONVS 13      /* torchvision vgg16.features idx 0..28 convs (DISTS_pt.py:36-49)      */
#define NQA_NUM_LEVELS 6      /* [x, relu1_2, relu2_2, relu3_3, relu4_3, relu5_3] (DISTS_pt.py:103)  */
#define NQA_NUM_CHANNELS 1475 /* sum(chns) = 3+64+128+256+512+512 (DISTS_pt.py:57)                   */

typedef struct nqa_ctx nqa_ctx;

typedef enum nqa_status {
  NQA_OK = 0,
  NQA_ERR_INVALID = -1, /* null pointer / bad enum / bad argument            */
  NQA_ERR_CUDA = -2,    /* CUDA runtime or driver failure, or no device      */
  NQA_ERR_SHAPE = -3,   /* B,H,W out of range (a pyramid level would vanish) */
  NQA_ERR_STATE = -4,   /* weights / alpha-beta not set yet                  */
  NQA_ERR_NOMEM = -5    /* workspace allocation failed                       */
} nqa_status;

/* Arithmetic of the VGG trunk (SURVEY.md §7 "parity tolerance"):
 *   FP32  — fp32 activations, fp32 CUDA-core FMA convolutions; per-frame scores within 1e-4
 *           relative of the reference's fp32 CPU result;
 *   BF16  — bf16 activations, tcgen05 (kind::f16, bf16 operands) with fp32 accumulation in TMEM;
 *           per-frame scores within 2e-3 absolute;
 *   BF16X3 — fp32 activations carried as bf16 hi+lo pairs, three tcgen05 MMAs per k-block
 *           (hi*hi + hi*lo + lo*hi), fp32 accumulation; fp32-class accuracy (1e-4 relative) on the
 *           tensor pipe.
 * Statistics, pooling and A-DISTS window math always accumulate in fp32 (final moments in fp64). */
typedef enum nqa_precision { NQA_PREC_FP32 = 0, NQA_PREC_BF16 = 1, NQA_PREC_BF16X3 = 2 } nqa_precision;

/* Message of the calling thread's most recent failing call ("" if none). */
NQA_API const char* nqa_last_error(void);
/* Library version, major*10000 + minor*100 + patch. */
NQA_API int nqa_version(void);

/* One context per GPU.  Replaces module construction + `.to(device)` (DISTS_pt.py:28-80,
 * ADISTS.py:36-69). */
NQA_API int nqa_ctx_create(int device, nqa_ctx** out);
NQA_API int nqa_ctx_destroy(nqa_ctx* ctx);

/* VGG16 conv parameters, OIHW fp32 weights and fp32 biases in torchvision features order
 * (state_dict keys stage{k}.{idx}.weight/bias; DISTS_pt.py:30-49).  `w[i]`/`b[i]` are "any"
 * pointers; the library repacks into its own layouts and keeps the copies. */
NQA_API int nqa_set_vgg_weights(nqa_ctx* ctx, const float* const w[NQA_NUM_CONVS], const float* const b[NQA_NUM_CONVS]);

/* DISTS channel weights alpha/beta, 1475 floats each ("any" pointers), RAW as stored in the module
 * (DISTS_pt.py:58-80); the division by w_sum = sum(alpha)+sum(beta) (DISTS_pt.py:127-129) happens
 * inside the scoring calls. */
NQA_API int nqa_set_alpha_beta(nqa_ctx* ctx, const float* alpha, const float* beta);

/* Channel weights used AS GIVEN: score = 1 - sum_c (a_c S1 + b_c S2), no division by the joint sum.  For the
 * reference's weight-normalisation variants, where the caller forms the effective weights itself and autograd
 * differentiates that transform: relu / w_sum.detach() / beta.detach() (DISTS_pt_original.py:111-121) and the
 * softmax over the 2950 logits (DISTS_pt_softmax.py:117-123).  d score_b / d a_c = -S1_bc, d score_b / d b_c = -S2_bc. */
NQA_API int nqa_set_channel_weights(nqa_ctx* ctx, const float* a, const float* b);

/* Stream-ordered forms of the two setters: the 2 x 5.9 KB are copied with cudaMemcpyAsync on `stream`, after the
 * work already enqueued there -- e.g. the torch ops that just produced alpha / beta (DISTS_pt.py:127-129 reads the
 * live parameters in every forward) -- and before scoring calls enqueued later on any stream of this context.  The
 * source must stay valid until the copy has run (device tensors: stream-ordered; pageable host memory: staged before
 * the call returns).  The synchronous forms above return after the copy has completed. */
NQA_API int nqa_set_alpha_beta_async(nqa_ctx* ctx, const float* alpha, const float* beta, void* stream);
NQA_API int nqa_set_channel_weights_async(nqa_ctx* ctx, const float* a, const float* b, void* stream);

/* DISTS.forward(x, y) (DISTS_pt.py:105-148): per-frame scores.
 *   x, y       device, [B,3,H,W] fp32
 *   score_out  device, [B] fp32
 *   s1s2_out   device or NULL, [B,2,1475] fp32: per-channel S1 then S2 (DISTS_pt.py:134,141), what
 *              the backward of the alpha/beta combination needs
 * Asynchronous on `stream`. */
NQA_API int nqa_dists_forward(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int precision,
                      float* score_out, float* s1s2_out, void* stream);

/* Same computation from HOST buffers: x_host/y_host/score_host are host pointers (pinned memory
 * gives full PCIe rate).  Frames are streamed host->device in micro-batches on a copy stream that
 * overlaps the compute stream; returns after the scores are in score_host.  This is the call
 * prep.py:186-190 makes implicitly with `.to(device)` ... `.cpu()`. */
NQA_API int nqa_dists_forward_host(nqa_ctx* ctx, const float* x_host, const float* y_host, int B, int H, int W,
                           int precision, float* score_host);

/* DISTS.forward_once(x) (DISTS_pt.py:91-103): levels 1..5 as NCHW fp32 device tensors
 * (level 0 is x itself).  feats_out[k-1] must hold B*C_k*H_k*W_k floats,
 * C = {64,128,256,512,512}, H_k = (H_{k-1}-1)/2+1 from level 2 on. */
NQA_API int nqa_dists_features(nqa_ctx* ctx, const float* x, int B, int H, int W, int precision,
                       float* const feats_out[NQA_NUM_LEVELS - 1], void* stream);

/* DISTS.forward_from_feats(feats0, feats1) (DISTS_pt.py:181-208): six NCHW fp32 device tensors per
 * image; level k has C_k channels and the pyramid sizes implied by H, W (level 0 = [B,3,H,W]). */
NQA_API int nqa_dists_from_feats(nqa_ctx* ctx, const float* const feats_x[NQA_NUM_LEVELS],
                         const float* const feats_y[NQA_NUM_LEVELS], int B, int H, int W, float* score_out,
                         float* s1s2_out, void* stream);

/* ADISTS.forward(x, y, as_loss=False) (ADISTS.py:137-197): per-frame scores [B] (device).
 * `window` is the gaussian window size (21 in the reference, ADISTS.py:34); a level smaller than
 * the window uses whole-map statistics (the reference's bare `except`, ADISTS.py:91-97,175-180).
 * Asymmetric: structure-probability maps and entropy weights come from x (ADISTS.py:147,153).
 * precision: NQA_PREC_FP32 or NQA_PREC_BF16 trunk; the A-DISTS statistics themselves are always fp32. */
NQA_API int nqa_adists_forward(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int window,
                       int precision, float* score_out, void* stream);

/* ADISTS.forward(as_map=True) (ADISTS.py:163,188-193): additionally writes the full-resolution map
 * 1 - sum_k bilinear(D_map_k -> (H, W), align_corners=False) to map_out [B,H,W] (device).  map_out may be NULL. */
NQA_API int nqa_adists_map(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int window, int precision,
                           float* score_out, float* map_out, void* stream);

/* nqa_adists_forward plus the intermediates of ADISTS.forward, for level-by-level parity checks against the reference
 * (all device pointers, each may be NULL):
 *   weights_out      [B,1475]  the clamped, renormalised entropy channel weights (ADISTS.py:127-135,152-160);
 *   level_terms_out  [B,6]     D_k = mean_hw sum_c w_c((1 - ps_k) T + ps_k S) per level; score = 1 - sum_k D_k
 *                              (ADISTS.py:165-191);
 *   ps_out[k]        [B, e_k]  the structure-probability map ps_x[k] of compute_prob (ADISTS.py:71-100), row-major
 *                              (h_k - window + 1) x (w_k - window + 1), or one value for a level smaller than the
 *                              window; e_k from nqa_adists_ps_elems. */
NQA_API int nqa_adists_parts(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int window, int precision,
                             float* score_out, float* weights_out, float* level_terms_out,
                             float* const ps_out[NQA_NUM_LEVELS], void* stream);
NQA_API int nqa_adists_ps_elems(int H, int W, int window, int level, size_t* elems);

/* DISTS and A-DISTS scores of the same batch from ONE trunk evaluation: what prep.py:186-190 /
 * test2_prep.py:151-155 compute with two model calls (four trunk forwards per pair in the reference).  Same trunk
 * activations as the single-metric calls; the DISTS scores equal nqa_dists_forward up to the summation order of the
 * fp32 moment partials (1e-6 relative: levels that travel as strip planes are summed by the strip-plane kernel; identical
 * for the fp32 / bf16x3 trunks), the A-DISTS scores equal nqa_adists_forward up to the rounding of the pooled
 * activations.  Both stay independent of the batch a frame travels in.  Needs VGG weights and alpha/beta. */
NQA_API int nqa_dists_adists_forward(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int window,
                                     int precision, float* dists_out, float* adists_out, void* stream);

/* Host-buffer forms (see nqa_dists_forward_host): the loop prep.py:181-190 runs per batch --
 * `adists_model(ref.to(device), render.to(device), as_loss=False)` and `dists_model(ref, render)` followed by
 * `.cpu()` -- with the frames streamed host -> device in micro-batches on a copy stream that overlaps the compute
 * stream.  x_host / y_host [B,3,H,W] and the score vectors [B] are host pointers (pinned memory gives the full PCIe
 * rate); the calls return when the scores are in host memory.  nqa_dists_adists_forward_host gives both metrics from
 * one trunk evaluation (needs alpha/beta). */
NQA_API int nqa_adists_forward_host(nqa_ctx* ctx, const float* x_host, const float* y_host, int B, int H, int W, int window,
                                    int precision, float* score_host);
NQA_API int nqa_dists_adists_forward_host(nqa_ctx* ctx, const float* x_host, const float* y_host, int B, int H, int W,
                                          int window, int precision, float* dists_host, float* adists_host);

/* The statistics half of ADISTS.forward (ADISTS.py:147-197) from six NCHW fp32 device tensors per
 * image (the reference's forward_once format, level 0 = the raw frames). */
NQA_API int nqa_adists_from_feats(nqa_ctx* ctx, const float* const feats_x[NQA_NUM_LEVELS],
                          const float* const feats_y[NQA_NUM_LEVELS], int B, int H, int W, int window,
                          float* score_out, void* stream);

/* Gradient of the DISTS scores w.r.t. the frames: autograd of DISTS.forward (DISTS_pt.py:105-148).
 * grad_score [B] (device) is d loss / d score_b; grad_x / grad_y [B,3,H,W] fp32 (device) receive d loss / d x, d y.
 * through_trunk = 1: require_grad=True semantics, the gradient flows through the VGG trunk (13 data-gradient
 * convolutions on the tensor pipe, L2pooling and statistics backward); through_trunk = 0: the reference's
 * default (trunk under no_grad), where only the level-0 statistics of the RAW frames are differentiable
 * (forward_once returns x itself, DISTS_pt.py:103).  The trunk forward is recomputed inside the call
 * (activation recomputation), so nothing has to be kept between nqa_dists_forward and this call.
 * precision: NQA_PREC_FP32 or NQA_PREC_BF16.  The alpha/beta gradients follow from the per-channel S1/S2
 * that nqa_dists_forward already returns. */
NQA_API int nqa_dists_backward(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int precision,
                               const float* grad_score, int through_trunk, float* grad_x, float* grad_y, void* stream);

/* The same forward / backward pair with the trunk activations KEPT between the two calls -- what torch autograd does
 * for DISTS.forward(require_grad=True) (DISTS_pt.py:105-148; the training loop run_final.py:168-229 calls backward right
 * after forward) -- instead of recomputing them: one trunk forward less per training step.
 * nqa_dists_forward_saved = nqa_dists_forward (scores, optional per-channel S1/S2) + the 13 conv outputs, 4 pooled
 * tensors and moments of every micro-batch stored in a library-owned tape buffer (142 MB per 512x512 image in bf16;
 * buffers are reused between steps).  *tape_out receives a handle, or 0 when the tape would exceed the limit
 * (NQA_TAPE_MAX_MB, default 65536) or cannot be allocated: the scores are still computed and the caller differentiates
 * with nqa_dists_backward.  nqa_dists_backward_saved consumes the tape (through_trunk = 1 semantics; x, y are the
 * same frames, needed for the level-0 term) and releases it; nqa_tape_release drops a tape that will not be
 * differentiated.  Forward and backward must be enqueued on the same stream.  precision: fp32 or bf16. */
NQA_API int nqa_dists_forward_saved(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int precision,
                                    float* score_out, float* s1s2_out, unsigned long long* tape_out, void* stream);
NQA_API int nqa_dists_backward_saved(nqa_ctx* ctx, unsigned long long tape, const float* x, const float* y,
                                     const float* grad_score, float* grad_x, float* grad_y, void* stream);
NQA_API int nqa_tape_release(nqa_ctx* ctx, unsigned long long tape);

/* Gradient of forward_from_feats (DISTS_pt.py:181-208) w.r.t. the 2 x 6 feature tensors (NCHW fp32, same shapes
 * as the inputs): what model_nr_v8.py:258-265 back-propagates into its decoder. */
NQA_API int nqa_dists_from_feats_backward(nqa_ctx* ctx, const float* const feats_x[NQA_NUM_LEVELS],
                                          const float* const feats_y[NQA_NUM_LEVELS], int B, int H, int W,
                                          const float* grad_score, float* const grad_fx[NQA_NUM_LEVELS],
                                          float* const grad_fy[NQA_NUM_LEVELS], void* stream);

/* Standalone L2pooling / Downsample on an NCHW fp32 device tensor (DISTS_pt.py:11-25,
 * ADISTS.py:16-31): out[B,C,(H-1)/2+1,(W-1)/2+1] = sqrt(hanning3x3(x^2, stride 2, pad 1) + 1e-12). */
NQA_API int nqa_l2pool(nqa_ctx* ctx, const float* x, int B, int C, int H, int W, float* out, void* stream);

/* Frame resizing in front of the scorers, on the device: x [B,C,H,W] fp32 -> out [B,C,OH,OW] fp32, bilinear with
 * align_corners=False.  antialias = 0 is F.interpolate(mode='bilinear', align_corners=False) as prep.py:91-96 and
 * the "pixel count" variant test2_prep.py:424-439 call it; antialias = 1 is the antialiased (triangle filter, support
 * = downscale factor) resize that torchvision's transforms.functional.resize performs on PIL images in
 * test2_prep.py:109-114 (short side 256) and :221-227 (256 x 256) -- equal to PIL up to its uint8 rounding. */
NQA_API int nqa_resize_bilinear(nqa_ctx* ctx, const float* x, int B, int C, int H, int W, int OH, int OW, int antialias,
                                float* out, void* stream);

/* Op-level 3x3 convolution + bias + ReLU (stride 1, zero pad 1) through the selected implementation,
 * for kernel parity tests: x [B,Cin,H,W], w [Cout,Cin,3,3], bias [Cout], out [B,Cout,H,W], all NCHW /
 * OIHW fp32 device tensors; Cin and Cout multiples of 64, Cout <= 512.  Equivalent of one
 * `Conv2d(3x3, padding=1)` + `ReLU` pair of torchvision's vgg16.features (DISTS_pt.py:36-49). */
NQA_API int nqa_conv3x3_relu(nqa_ctx* ctx, const float* x, const float* w, const float* bias, int B, int Cin, int Cout,
                             int H, int W, int precision, float* out, void* stream);

/* Bytes of device workspace the context currently holds (after warm-up this stops growing). */
NQA_API int nqa_workspace_bytes(nqa_ctx* ctx, size_t* bytes);

/* Number of kernels this library has launched on the context since creation (bench accounting). */
NQA_API int nqa_launch_count(nqa_ctx* ctx, uint64_t* launches);

/* Small-frame calls (B * H * W <= 2^21, e.g. the reference's 256 x 256 frames at batch 4, nerf_qa/settings.py:1 and
 * prep.py:163) of nqa_dists_forward (without s1s2_out), nqa_adists_forward and nqa_dists_adists_forward are replayed
 * from a CUDA graph captured at the second call of a shape (NQA_GRAPHS=0 disables it).  Number of replays so far: */
NQA_API int nqa_graph_replays(nqa_ctx* ctx, uint64_t* replays);

/* Per-kernel-family device timing with CUDA events recorded on the launching stream around every
 * launch (bench.py's live roofline numbers).  nqa_profile_read synchronises the device, returns the
 * milliseconds and launch counts accumulated since the previous read, and resets them. */
#define NQA_NUM_KERNEL_FAMILIES 12
enum nqa_kernel_family {
  NQA_K_CONV_TC = 0,  /* tcgen05 implicit-GEMM 3x3 conv + bias + ReLU          */
  NQA_K_CONV1_1 = 1,  /* normalise + conv1_1 (CUDA cores)                      */
  NQA_K_CONV_F32 = 2, /* strict fp32 conv (CUDA cores)                         */
  NQA_K_L2POOL = 3,   /* L2pooling                                             */
  NQA_K_STATS = 4,    /* DISTS per-level statistics (NHWC + planar level 0)    */
  NQA_K_FINALIZE = 5, /* S1/S2 + alpha/beta combination                        */
  NQA_K_LAYOUT = 6,   /* NHWC <-> NCHW conversions                             */
  NQA_K_ADISTS = 7,   /* A-DISTS: channel sums, entropy weights, ps chain      */
  NQA_K_POOL_STATS = 8, /* fused L2pooling + DISTS statistics (levels 1..4)    */
  NQA_K_STAGE1 = 9,   /* fused normalise + conv1_1 + conv1_2 (tcgen05)         */
  NQA_K_BACKWARD = 10, /* statistics / pooling / conv1_1 gradient kernels      */
  NQA_K_ADISTS_WINDOW = 11 /* A-DISTS: the local-window statistics kernels     */
};
NQA_API int nqa_profile_enable(nqa_ctx* ctx, int on);
NQA_API int nqa_profile_read(nqa_ctx* ctx, double ms[NQA_NUM_KERNEL_FAMILIES], uint64_t launches[NQA_NUM_KERNEL_FAMILIES]);

/* Measured arithmetic rates of this GPU for the roofline denominators MEASURED_PEAKS.json does not carry (it has the
 * HBM copy and the cuBLAS bf16 figures): a short register-resident micro-benchmark on every SM, best of three, in
 * multiply-adds per second.  Synchronises the device. */
enum nqa_rate {
  NQA_RATE_FP32_FMA = 0,       /* packed fp32 FMAs on the CUDA cores (what the fp32 window kernel issues) */
  NQA_RATE_MMA_SYNC_BF16 = 1   /* warp-level mma.sync m16n8k16 bf16 -> fp32 (the banded-Toeplitz window kernel) */
};
NQA_API int nqa_measure_rate(nqa_ctx* ctx, int which, double* fma_per_s);

/* Micro-batch (frame pairs per trunk pass) the scheduler would use for this shape. */
NQA_API int nqa_micro_batch(nqa_ctx* ctx, int B, int H, int W, int precision, int* pairs);

#ifdef __cplusplus
}
#endif
#endif /* NERFQA_B200_H */
