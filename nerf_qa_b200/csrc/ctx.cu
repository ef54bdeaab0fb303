// C-ABI of nerfqa_b200 (include/nerfqa_b200.h): context, weight packing, workspace arena, micro-batch
// scheduling and the per-call kernel sequences for DISTS / A-DISTS.
//
// Reference call structure being replaced: DISTS.forward -> forward_once x2 -> statistics loop
// (nerf_qa/DISTS_pytorch/DISTS_pt.py:91-148), ADISTS.forward (nerf_qa/ADISTS/ADISTS.py:137-197).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <string>
#include <type_traits>
#include <vector>

#include "../../include/nerfqa_b200.h"
#include "adists.h"
#include "common.cuh"
#include "kernels.h"

using namespace nqa;

static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(NQA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static const int kConvCin[kNumConvs] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512};
static const int kConvCout[kNumConvs] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
static const int kStageConvs[5] = {2, 2, 3, 3, 3};

struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  void reset() { used = 0; }
  void* take(size_t bytes) {
    const size_t a = (used + 255) & ~(size_t)255;
    used = a + bytes;
    return base + a;
  }
};

struct TapeBuf {
  char* base = nullptr;
  size_t cap = 0;
  bool in_use = false;
};
// Activations a backward walk needs: the 13 conv outputs, the 4 pooled tensors and the per-channel moments of one
// micro-batch.  They live either in the workspace arena (activation recomputation inside nqa_dists_backward) or in a
// tape buffer that survives from nqa_dists_forward_saved to nqa_dists_backward_saved.
struct SavedActs {
  void* act[kNumConvs] = {};
  void* pool[4] = {};
  void* act_lo[kNumConvs] = {};   // bf16x3: residual carriers of the same tensors
  void* pool_lo[4] = {};
  float* moments = nullptr;
  int pairs = 0;
};
struct Tape {
  unsigned long long id = 0;
  int buf = -1;
  int B = 0, H = 0, W = 0, prec = 0, mb = 0;
  std::vector<SavedActs> chunks;   // one per micro-batch of mb pairs
};

struct nqa_ctx {
  int device = 0;
  int sm_count = kNumSMs;
  bool have_weights = false, have_alpha_beta = false;
  bool ab_normalize = true;              // divide alpha/beta by their joint sum inside the kernels (DISTS_pt.py:127-129)
  float* w11 = nullptr;                  // conv1_1 [27][64] (CUDA-core path)
  bf16* w11_img = nullptr;               // conv1_1 [2][64*32] UMMA operand image (tensor path)
  float* bias[kNumConvs] = {};           // fp32
  float* w_f32[kNumConvs] = {};          // [9][cin][cout]
  bf16* w_hi[kNumConvs] = {};            // [cout][9][cin]
  bf16* w_lo[kNumConvs] = {};            // residual of the bf16 rounding (bf16x3)
  float* w_bwd_f32[kNumConvs] = {};      // data-gradient weights [9][cout][cin] (rotated taps)
  bf16* w_bwd_hi[kNumConvs] = {};        // data-gradient weights [cin][9][cout] bf16
  bf16* w_bwd_lo[kNumConvs] = {};        // their bf16 rounding residual (bf16x3 backward)
  float* w11_bwd = nullptr;              // conv1_1 backward [9][64][3] (device staging of the packed weights)
  Conv11BwdWeights w11_bwd_host;         // the same, passed to the backward kernel as a parameter (constant bank)
  bf16* w11_bwd_hi = nullptr;            // conv1_1 data-gradient weights for the tensor pipe: [64 (3 used)][9][64] bf16
  std::vector<TapeBuf> tape_bufs;        // device buffers of the activation tapes (reused between training steps)
  std::vector<Tape> tapes;               // live tapes
  unsigned long long next_tape = 1;
  size_t tape_limit = (size_t)64 << 30;  // NQA_TAPE_MAX_MB: above this the saved-activation forward declines (tape 0)
  bool conv11_bwd_tc = true;             // NQA_CONV11_BWD_TC=0: CUDA-core conv1_1 backward in the bf16 trunk too
  float* alpha = nullptr;
  float* beta = nullptr;
  CUtensorMap w12_map;                   // conv1_2 weights for the fused stage-1 kernel, box {64, 64}
  CUtensorMap w12_map_pair;              // the same with box {64, 32}: CTA-pair version (each CTA holds half the channels)
  bool stage1_pair = true;               // NQA_STAGE1_PAIR=0: fused stage 1 on single CTAs instead of CTA pairs
  bool stage1_a2_linear = true;          // NQA_STAGE1_A2_LINEAR=0: relu1_1 tile in the 128-byte-swizzled layout
  bool fuse_stage1 = true;               // NQA_STAGE1_FUSED=0 falls back to conv1_1 + conv1_2 kernels
  int micro_images = 16;                 // full-HD-equivalent images per trunk pass (NQA_MICRO_IMAGES)
  bool fuse_pool_stats = true;           // NQA_POOL_STATS_FUSED=0 falls back to separate pooling / statistics kernels
  bool pool_stats_tma = true;            // NQA_POOL_STATS_TMA=0: register-fed pooling + moments kernel for bf16 levels too
  AdistsOptions adists_opt;              // NQA_ADISTS_MMA / NQA_ADISTS_MMA_TERMS
  bool adists_fused_sums = true;         // NQA_ADISTS_FUSED_SUMS=0: separate planes / sums / entropy kernels (bf16 trunk)
  int adists_strips = 2;                 // NQA_ADISTS_STRIPS: 2 = the last convolution of a stage writes strip planes and one
                                         // kernel pools + sums them (A-DISTS-only calls of the bf16 trunk); 1 = strip planes
                                         // from the NHWC level (validation); 0 = planar planes
  ConvTcConfig conv_cfg;                 // kernel selection of the tensor-pipe convolutions (environment read at creation)
  // Encoded tensor maps of the convolution launches, keyed on (buffers, shape): the arena hands out the same addresses
  // for the same shape, so after the first call at a shape no launch goes through cuTensorMapEncodeTiled again.
  struct PlanKey {
    const void *in, *in_lo, *wk, *out, *out_lo;
    int n, h, w, cin, cout, split;
    bool operator==(const PlanKey& o) const {
      return in == o.in && in_lo == o.in_lo && wk == o.wk && out == o.out && out_lo == o.out_lo && n == o.n && h == o.h &&
             w == o.w && cin == o.cin && cout == o.cout && split == o.split;
    }
  };
  std::vector<std::pair<PlanKey, ConvTcPlan>> plan_cache;
  uint64_t plan_hits = 0, plan_misses = 0;
  // Cross-stream ordering of the context's shared state (workspace arena, channel weights, staging buffers): every
  // entry point records ev_last on its stream when it has enqueued its work, and an entry point that runs on a
  // DIFFERENT stream than the previous one first makes its stream wait for that event.
  cudaEvent_t ev_last = nullptr;
  cudaStream_t last_stream = nullptr;
  bool have_last = false;
  // CUDA graphs of the device entry points for small frames (the reference's own call pattern is 256 x 256 at batch 4,
  // nerf_qa/settings.py:1, prep.py:163): ~60 launches of a few microseconds each are launch-bound, so after one plain
  // call at a shape the sequence is captured once -- on staging buffers, so the graph does not depend on the caller's
  // addresses -- and replayed.  NQA_GRAPHS=0 switches this off.
  struct GraphEntry {
    int kind, B, H, W, prec, window, ab_norm;
    int state = 0;                 // 0 = seen once (warmed up), 2 = captured, -1 = capture failed (plain path)
    cudaGraphExec_t exec = nullptr;
    float *sx = nullptr, *sy = nullptr, *sout = nullptr;
    int n_out = 1;
    uint64_t launches = 0;
  };
  std::vector<GraphEntry> graphs;
  bool use_graphs = true;
  long long graph_max_pixels = 1ll << 21;   // B * H * W up to 32 frames of 256 x 256
  uint64_t graph_replays = 0;
  cudaEvent_t ev_graph_in = nullptr, ev_graph_out = nullptr;
  Arena arena;
  uint64_t launches = 0;
  cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
  // Optional second lane (NQA_LANES=2): consecutive micro-batches alternate between two internal streams so that the
  // HBM-bound pooling / statistics kernels of one micro-batch share the SMs with the tensor-bound convolutions of
  // the other.  Measured on B200: no gain (428 vs 430 pairs/s) -- the step runs at the 1000 W power cap, the overlap
  // only lowers the SM clock (1556 vs 1610 MHz): the path is energy-bound, not latency-bound.  Default: one lane.
  int lanes = 1;
  cudaStream_t lane_stream[2] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {};
  cudaEvent_t ev_copied[2] = {}, ev_consumed[2] = {};
  float* stage_in[2] = {};               // device staging for the host path
  size_t stage_cap = 0;
  // optional per-family profiling (CUDA events on the launching stream)
  bool profiling = false;
  struct ProfRec { int family; cudaEvent_t a, b; uint64_t launches; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
};

// Brackets the launches made inside its scope with two events when profiling is on.
struct ProfScope {
  nqa_ctx* ctx;
  cudaStream_t st;
  int family;
  uint64_t before;
  cudaEvent_t a = nullptr;
  static cudaEvent_t get(nqa_ctx* c) {
    if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  ProfScope(nqa_ctx* c, int fam, cudaStream_t s) : ctx(c), st(s), family(fam), before(c->launches) {
    if (ctx->profiling) { a = get(ctx); cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (a) {
      cudaEvent_t b = get(ctx);
      cudaEventRecord(b, st);
      ctx->prof_recs.push_back({family, a, b, ctx->launches - before});
    }
  }
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

void drop_graphs(nqa_ctx* ctx);

int ensure_arena(nqa_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->arena.cap) return NQA_OK;
  if (ctx->arena.base) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaFree(ctx->arena.base));
    ctx->arena.base = nullptr;
    ctx->arena.cap = 0;
    ctx->plan_cache.clear();   // the cached tensor maps point into the old arena
    drop_graphs(ctx);          // and so do the captured launch sequences
  }
  const size_t want = bytes + (bytes >> 4) + (1 << 20);
  void* p = nullptr;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    return fail(NQA_ERR_NOMEM, "workspace allocation of %zu bytes failed", want);
  }
  ctx->arena.base = static_cast<char*>(p);
  ctx->arena.cap = want;
  return NQA_OK;
}

size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// Orders a call on `st` after whatever the previous call enqueued on another stream (see nqa_ctx::ev_last).
int enter_stream(nqa_ctx* ctx, cudaStream_t st) {
  if (ctx->have_last && ctx->last_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_last, 0));
  return NQA_OK;
}
int leave_stream(nqa_ctx* ctx, cudaStream_t st) {
  CUDA_TRY(cudaEventRecord(ctx->ev_last, st));
  ctx->last_stream = st;
  ctx->have_last = true;
  return NQA_OK;
}
// RAII form for the entry points: waits on entry, records on every exit path.
struct StreamOrder {
  nqa_ctx* ctx;
  cudaStream_t st;
  int rc;
  StreamOrder(nqa_ctx* c, cudaStream_t s) : ctx(c), st(s), rc(enter_stream(c, s)) {}
  ~StreamOrder() { leave_stream(ctx, st); }
};

// conv_tc_make_plan through the context's cache of encoded tensor maps
int cached_plan(nqa_ctx* ctx, ConvTcPlan* plan, const bf16* in, const bf16* in_lo, const bf16* wk, const bf16* wk_lo,
                const float* bias, bf16* out, bf16* out_lo, int n, int h, int w, int cin, int cout, bool split,
                const char** err, bool strips_out = false, bool split_on_pairs = false) {
  const nqa_ctx::PlanKey key{in, in_lo, wk, out, out_lo, n, h, w, cin, cout,
                             (split ? 1 : 0) | (strips_out ? 2 : 0) | (split_on_pairs ? 4 : 0)};
  for (auto& e : ctx->plan_cache)
    if (e.first == key) {
      *plan = e.second;
      plan->bias = bias;
      ++ctx->plan_hits;
      return 0;
    }
  const int rc = conv_tc_make_plan(plan, ctx->conv_cfg, in, in_lo, wk, wk_lo, bias, out, out_lo, n, h, w, cin, cout, split, err,
                                   strips_out, split_on_pairs);
  if (rc) return rc;
  ++ctx->plan_misses;
  if (ctx->plan_cache.size() >= 512) ctx->plan_cache.clear();
  ctx->plan_cache.emplace_back(key, *plan);
  return 0;
}

void level_dims(int H, int W, int (&h)[kNumLevels], int (&w)[kNumLevels]) {
  h[0] = h[1] = H;
  w[0] = w[1] = W;
  for (int k = 2; k < kNumLevels; ++k) {
    h[k] = pooled(h[k - 1]);
    w[k] = pooled(w[k - 1]);
  }
}

// Largest activation of the trunk per image, in elements.  For ordinary frames that is relu1_x (64 x H x W); once a
// pyramid dimension is stuck at 1 the channel doubling wins (a 1x1 frame has 512 values at levels 4 and 5).
size_t max_act_elems(int H, int W) {
  int h[kNumLevels], w[kNumLevels];
  level_dims(H, W, h, w);
  size_t m = 0;
  for (int k = 1; k < kNumLevels; ++k) {
    const size_t e = (size_t)level_channels(k) * h[k] * w[k];
    if (e > m) m = e;
  }
  return m;
}

int check_shape(int B, int H, int W) {
  if (B < 1) return fail(NQA_ERR_SHAPE, "batch size %d < 1", B);
  if (H < 1 || W < 1) return fail(NQA_ERR_SHAPE, "image size %dx%d is empty", H, W);
  if ((long long)H * W > (1ll << 28)) return fail(NQA_ERR_SHAPE, "image size %dx%d too large", H, W);
  return NQA_OK;
}

int elem_bytes(int prec) { return prec == NQA_PREC_FP32 ? 4 : 2; }
int carriers(int prec) { return prec == NQA_PREC_BF16X3 ? 2 : 1; }

// images per trunk pass: ~16 full-HD frames keep every layer's persistent grid at many tiles per CTA (small tail
// effects on the 68 x 120 level) while the ping-pong activations stay below 10 GB (NQA_MICRO_IMAGES overrides).
int pick_micro_pairs(const nqa_ctx* ctx, int B, int H, int W) {
  const long long target = (long long)ctx->micro_images * 1080 * 1920;
  long long pairs = target / (2ll * H * W);
  if (pairs < 1) pairs = 1;
  if (pairs > 256) pairs = 256;
  if (pairs > B) pairs = B;
  return (int)pairs;
}

void drop_graphs(nqa_ctx* ctx) {
  for (auto& g : ctx->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    cudaFree(g.sx);
    cudaFree(g.sy);
    cudaFree(g.sout);
  }
  ctx->graphs.clear();
}

// Runs `body(x, y, outs, stream)` -- the plain launch sequence of an entry point writing n_out vectors of B scores --
// either directly on `st` or, for small frames from the second call at a shape on, as a replay of its captured graph.
template <typename F>
int graphed_call(nqa_ctx* ctx, int kind, const float* x, const float* y, int B, int H, int W, int prec, int window,
                 float* const* outs, int n_out, cudaStream_t st, F&& body) {
  const bool eligible = ctx->use_graphs && !ctx->profiling && ctx->lanes == 1 && (long long)B * H * W <= ctx->graph_max_pixels;
  if (!eligible) return body(x, y, outs, st);
  nqa_ctx::GraphEntry* e = nullptr;
  for (auto& g : ctx->graphs)
    if (g.kind == kind && g.B == B && g.H == H && g.W == W && g.prec == prec && g.window == window &&
        g.ab_norm == (ctx->ab_normalize ? 1 : 0)) {
      e = &g;
      break;
    }
  if (!e) {   // first call at this shape: plain (grows the arena, sets kernel attributes, fills the tensor-map cache)
    const int rc = body(x, y, outs, st);
    if (rc == NQA_OK) {
      if (ctx->graphs.size() >= 64) drop_graphs(ctx);
      nqa_ctx::GraphEntry g;
      g.kind = kind; g.B = B; g.H = H; g.W = W; g.prec = prec; g.window = window; g.ab_norm = ctx->ab_normalize ? 1 : 0;
      g.n_out = n_out;
      ctx->graphs.push_back(g);
    }
    return rc;
  }
  if (e->state < 0) return body(x, y, outs, st);
  const size_t frame_bytes = (size_t)B * 3 * H * W * sizeof(float);
  cudaStream_t gs = ctx->compute_stream;
  if (e->state == 0) {   // second call: capture on the staging buffers
    e->state = -1;
    if (cudaMalloc(&e->sx, frame_bytes) != cudaSuccess || cudaMalloc(&e->sy, frame_bytes) != cudaSuccess ||
        cudaMalloc(&e->sout, (size_t)n_out * B * sizeof(float)) != cudaSuccess) {
      cudaGetLastError();
      return body(x, y, outs, st);
    }
    float* souts[2] = {e->sout, e->sout + B};
    const uint64_t before = ctx->launches;
    cudaGraph_t graph = nullptr;
    // everything earlier on the internal stream has to be done before it starts capturing
    CUDA_TRY(cudaStreamSynchronize(gs));
    if (cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return body(x, y, outs, st);
    }
    const int rc = body(e->sx, e->sy, souts, gs);
    const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
    e->launches = ctx->launches - before;
    ctx->launches = before;   // nothing ran yet
    if (rc != NQA_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&e->exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      e->exec = nullptr;
      return body(x, y, outs, st);
    }
    cudaGraphDestroy(graph);
    e->state = 2;
  }
  // replay: internal stream after the caller's, the caller's after the internal one
  CUDA_TRY(cudaEventRecord(ctx->ev_graph_in, st));
  CUDA_TRY(cudaStreamWaitEvent(gs, ctx->ev_graph_in, 0));
  CUDA_TRY(cudaMemcpyAsync(e->sx, x, frame_bytes, cudaMemcpyDeviceToDevice, gs));
  CUDA_TRY(cudaMemcpyAsync(e->sy, y, frame_bytes, cudaMemcpyDeviceToDevice, gs));
  CUDA_TRY(cudaGraphLaunch(e->exec, gs));
  for (int o = 0; o < n_out; ++o)
    CUDA_TRY(cudaMemcpyAsync(outs[o], e->sout + (size_t)o * B, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, gs));
  CUDA_TRY(cudaEventRecord(ctx->ev_graph_out, gs));
  CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_graph_out, 0));
  ctx->launches += e->launches;
  ++ctx->graph_replays;
  return NQA_OK;
}

// ---- weight packing kernels ---------------------------------------------------------------------
__global__ void pack_w11_kernel(const float* __restrict__ w, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 27*64
  if (i < 27 * 64) {
    const int k = i / 64, co = i % 64;
    out[i] = w[co * 27 + k];
  }
}
__global__ void pack_w_kernel(const float* __restrict__ w, int cin, int cout, float* __restrict__ w_f32,
                              bf16* __restrict__ w_hi, bf16* __restrict__ w_lo) {
  const size_t total = (size_t)cout * cin * 9;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9);
    const int ci = (int)((i / 9) % cin);
    const int co = (int)(i / (9 * (size_t)cin));
    const float v = w[i];  // OIHW
    w_f32[((size_t)tap * cin + ci) * cout + co] = v;
    const bf16 hi = __float2bfloat16_rn(v);
    const size_t kidx = (size_t)co * 9 * cin + (size_t)tap * cin + ci;
    w_hi[kidx] = hi;
    w_lo[kidx] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// ---- trunk --------------------------------------------------------------------------------------
template <typename T>
struct TrunkBufs {
  T* ping[2] = {nullptr, nullptr};
  T* ping_lo[2] = {nullptr, nullptr};   // bf16x3 residual carrier
};

template <typename T>
size_t trunk_bytes(int n, int H, int W, int prec) {
  const size_t big = align256((size_t)n * max_act_elems(H, W) * sizeof(T));
  return 2 * big * carriers(prec) + 4096;
}

template <typename T>
void trunk_alloc(nqa_ctx* ctx, TrunkBufs<T>* b, int n, int H, int W, int prec) {
  const size_t big = (size_t)n * max_act_elems(H, W) * sizeof(T);
  for (int i = 0; i < 2; ++i) {
    b->ping[i] = static_cast<T*>(ctx->arena.take(big));
    if (prec == NQA_PREC_BF16X3) b->ping_lo[i] = static_cast<T*>(ctx->arena.take(big));
  }
}

// strips_out (bf16 trunk, CTA-pair layers): `out` is a strip-plane buffer [n][cout][w/8][h][8] instead of NHWC
template <typename T>
int conv_layer(nqa_ctx* ctx, int li, const T* in, const T* in_lo, T* out, T* out_lo, int n, int h, int w, int prec,
               cudaStream_t st, bool strips_out = false);

template <typename T>
void conv1_1_layer(nqa_ctx* ctx, const float* x, const float* y, int nx, int n, int h, int w, T* out, T* out_lo,
                   cudaStream_t st);
template <>
void conv1_1_layer<float>(nqa_ctx* ctx, const float* x, const float* y, int nx, int n, int h, int w, float* out,
                          float* out_lo, cudaStream_t st) {
  ProfScope ps(ctx, NQA_K_CONV1_1, st);
  ctx->launches += launch_conv1_1<float>(x, y, nx, n, h, w, ctx->w11, ctx->bias[0], out, out_lo, st);
}
template <>
void conv1_1_layer<bf16>(nqa_ctx* ctx, const float* x, const float* y, int nx, int n, int h, int w, bf16* out,
                         bf16* out_lo, cudaStream_t st) {
  ProfScope ps(ctx, NQA_K_CONV1_1, st);
  ctx->launches += launch_conv1_1_tc(x, y, nx, n, h, w, ctx->w11_img, ctx->bias[0], out, out_lo, ctx->sm_count, st);
}

template <>
int conv_layer<float>(nqa_ctx* ctx, int li, const float* in, const float*, float* out, float*, int n, int h, int w, int,
                      cudaStream_t st, bool) {
  ProfScope ps(ctx, NQA_K_CONV_F32, st);
  ctx->launches += launch_conv3x3_f32(in, out, ctx->w_f32[li], ctx->bias[li], n, h, w, kConvCin[li], kConvCout[li], st);
  return NQA_OK;
}
template <>
int conv_layer<bf16>(nqa_ctx* ctx, int li, const bf16* in, const bf16* in_lo, bf16* out, bf16* out_lo, int n, int h,
                     int w, int prec, cudaStream_t st, bool strips_out) {
  ConvTcPlan plan;
  const char* err = "";
  const bool split = prec == NQA_PREC_BF16X3;
  if (cached_plan(ctx, &plan, in, in_lo, ctx->w_hi[li], ctx->w_lo[li], ctx->bias[li], out, out_lo, n, h, w,
                  kConvCin[li], kConvCout[li], split, &err, strips_out, /*split_on_pairs=*/split))
    return fail(NQA_ERR_CUDA, "%s (layer %d, %dx%dx%d)", err, li, n, h, w);
  ProfScope ps(ctx, NQA_K_CONV_TC, st);
  const int nl = launch_conv3x3_tc(plan, st);
  if (nl <= 0) return fail(NQA_ERR_CUDA, "convolution of layer %d could not be launched (%dx%dx%d)", li, n, h, w);
  ctx->launches += nl;
  return NQA_OK;
}

// Runs normalise + VGG16 on images [x(0..nx), y(0..n-nx)], ping-ponging between two NHWC buffers sized
// for the largest activation (64 x H x W).  After the last conv of stage k (1..5) calls
// on_level(k, feat, feat_lo, h, w, pooled_dst) with the stage's NHWC output enqueued; pooled_dst is the buffer the
// next stage reads (nullptr after stage 5).  A consumer that already wrote the L2-pooled features there (the fused
// pool + statistics kernel) returns true; otherwise the trunk enqueues the pooling kernel itself.  The level buffer
// stays valid until the next stage's first conv is enqueued, i.e. consumers enqueued inside on_level are safe.
// strips_dst (optional, bf16 trunk): strips_dst[k] != nullptr makes the last convolution of stage k write its output as
// strip planes into that buffer (trunk_strips_capable says whether it can) instead of the NHWC ping-pong slot; on_level
// is then called with feat == nullptr and must do the pooling itself (return true) if pooled_dst is set.
template <typename T, typename F>
int run_trunk(nqa_ctx* ctx, const float* x, const float* y, int nx, int n, int H, int W, int prec, TrunkBufs<T>& bufs,
              cudaStream_t st, F&& on_level, bf16* const* strips_dst = nullptr) {
  const bool split = prec == NQA_PREC_BF16X3;
  int h = H, w = W;
  int cur = 0;  // ping-pong slot holding the live activation
  int li = 0;
  for (int stage = 0; stage < 5; ++stage) {
    bool level_in_strips = false;
    for (int j = 0; j < kStageConvs[stage]; ++j, ++li) {
      if (li == 0 && std::is_same<T, bf16>::value && !split && ctx->fuse_stage1) {
        // normalise + conv1_1 + conv1_2 in one kernel; relu1_1 never reaches HBM
        ProfScope ps(ctx, NQA_K_STAGE1, st);
        const bool pair = ctx->stage1_pair && (ctx->sm_count & 1) == 0;
        const int nl = launch_stage1_fused(x, y, nx, n, h, w, ctx->w11_img, ctx->bias[0], pair ? ctx->w12_map_pair : ctx->w12_map,
                                           ctx->bias[1], strips_dst && strips_dst[1] ? strips_dst[1] : reinterpret_cast<bf16*>(bufs.ping[1]),
                                           ctx->sm_count, pair, ctx->stage1_a2_linear, st, strips_dst && strips_dst[1]);
        if (nl <= 0) return fail(NQA_ERR_CUDA, "stage-1 kernel: tensor map or cluster launch failed (%dx%dx%d)", n, h, w);
        ctx->launches += nl;
        cur = 1;
        ++li;
        ++j;
      } else if (li == 0) {
        conv1_1_layer<T>(ctx, x, y, nx, n, h, w, bufs.ping[0], split ? bufs.ping_lo[0] : nullptr, st);
        cur = 0;
      } else if (strips_dst && strips_dst[stage + 1] && j == kStageConvs[stage] - 1 && li > 1) {
        // last convolution of the stage -> strip planes; the ping-pong slot it would have written stays free for the
        // pooled tensor
        const int rc = conv_layer<T>(ctx, li, bufs.ping[cur], nullptr, reinterpret_cast<T*>(strips_dst[stage + 1]), nullptr, n, h, w,
                                     prec, st, true);
        if (rc != NQA_OK) return rc;
        level_in_strips = true;
      } else {
        const int rc = conv_layer<T>(ctx, li, bufs.ping[cur], split ? bufs.ping_lo[cur] : nullptr, bufs.ping[cur ^ 1],
                                     split ? bufs.ping_lo[cur ^ 1] : nullptr, n, h, w, prec, st);
        if (rc != NQA_OK) return rc;
        cur ^= 1;
      }
    }
    if (stage == 0 && strips_dst && strips_dst[1] && std::is_same<T, bf16>::value && !split && ctx->fuse_stage1) level_in_strips = true;
    T* pooled_dst = stage < 4 ? bufs.ping[cur ^ 1] : nullptr;
    const bool pooled_done = on_level(stage + 1, level_in_strips ? nullptr : bufs.ping[cur], split ? bufs.ping_lo[cur] : nullptr, h, w,
                                      pooled_dst);
    if (level_in_strips && stage < 4 && !pooled_done) return fail(NQA_ERR_INVALID, "strip-plane level %d was not pooled", stage + 1);
    if (stage < 4) {
      if (!pooled_done) {
        ProfScope ps(ctx, NQA_K_L2POOL, st);
        const int nl = launch_l2pool_nhwc<T>(bufs.ping[cur], split ? bufs.ping_lo[cur] : nullptr, bufs.ping[cur ^ 1],
                                             split ? bufs.ping_lo[cur ^ 1] : nullptr, n, h, w, kConvCout[li - 1], st);
        if (nl <= 0) return fail(NQA_ERR_SHAPE, "L2pooling: %d images of %dx%d exceed the launch grid", n, h, w);
        ctx->launches += nl;
      }
      h = pooled(h);
      w = pooled(w);
      cur ^= 1;
    }
  }
  return NQA_OK;
}

// Whether the kernel that produces level k (1..5) of the bf16 trunk can write strip planes: the fused stage-1 kernel for
// level 1, a CTA-pair convolution for the others (the default kernel selection; switches that pick other kernels turn
// the strip path off for that level).
bool trunk_strips_capable(const nqa_ctx* ctx, int k) {
  static const int last_conv[kNumLevels] = {-1, 1, 3, 6, 9, 12};
  if (k == 1) return ctx->fuse_stage1;
  return k >= 2 && k < kNumLevels && conv_tc_uses_pair(ctx->conv_cfg, kConvCin[last_conv[k]], kConvCout[last_conv[k]], false);
}

template <typename T>
void level_to_nchw(nqa_ctx* ctx, const T* f, const T* f_lo, float* dst, int n, int c, int h, int w, cudaStream_t st);
template <>
void level_to_nchw<float>(nqa_ctx* ctx, const float* f, const float*, float* dst, int n, int c, int h, int w,
                          cudaStream_t st) {
  ProfScope ps(ctx, NQA_K_LAYOUT, st);
  ctx->launches += launch_nhwc_to_nchw<float>(f, dst, n, c, h, w, st);
}
template <>
void level_to_nchw<bf16>(nqa_ctx* ctx, const bf16* f, const bf16* f_lo, float* dst, int n, int c, int h, int w,
                         cudaStream_t st) {
  ProfScope ps(ctx, NQA_K_LAYOUT, st);
  if (f_lo)
    ctx->launches += launch_nhwc_split_to_nchw(f, f_lo, dst, n, c, h, w, st);
  else
    ctx->launches += launch_nhwc_to_nchw<bf16>(f, dst, n, c, h, w, st);
}

// ---- DISTS: one micro-batch --------------------------------------------------------------------
struct PartialBufs {
  float* p[kNumLevels];
  int chunks[kNumLevels];
  double* terms;
};

// strips_mask: levels whose moments come from the strip-plane kernel (adists_strips.cu) and follow its chunking
size_t partial_bytes(int pairs, int H, int W, int prec, PartialBufs* pb /* may be null */, bool planar_all,
                     bool fused_pool = false, unsigned strips_mask = 0) {
  int h[kNumLevels], w[kNumLevels];
  level_dims(H, W, h, w);
  size_t total = 0;
  for (int k = 0; k < kNumLevels; ++k) {
    const int c = level_channels(k);
    const int chunks = ((strips_mask >> k) & 1) ? adists_strips_chunks(c, h[k], w[k])
                       : (k == 0 || planar_all) ? stats_chunks_planar(h[k] * w[k], c, pairs)
                       : (fused_pool && k < kNumLevels - 1) ? pool_stats_chunks(h[k], w[k], c, elem_bytes(prec))
                                                            : stats_chunks_nhwc(h[k] * w[k], c, elem_bytes(prec), pairs);
    if (pb) pb->chunks[k] = chunks;
    total += align256(((size_t)pairs * chunks * 5 * c + (size_t)pairs * 2 * c) * sizeof(float));   // partial sums + pivots
  }
  total += align256((size_t)pairs * kTotalChannels * sizeof(double));
  return total + 4096;
}

void partial_alloc(nqa_ctx* ctx, int pairs, PartialBufs* pb) {
  for (int k = 0; k < kNumLevels; ++k)
    pb->p[k] = static_cast<float*>(ctx->arena.take(((size_t)pairs * pb->chunks[k] * 5 + (size_t)pairs * 2) * level_channels(k) * sizeof(float)));
  pb->terms = static_cast<double*>(ctx->arena.take((size_t)pairs * kTotalChannels * sizeof(double)));
}

int finalize(nqa_ctx* ctx, const PartialBufs& pb, int pairs, int H, int W, float* score, float* s1s2, cudaStream_t st,
             float* moments = nullptr, int levels_valid = kNumLevels) {
  int h[kNumLevels], w[kNumLevels];
  level_dims(H, W, h, w);
  FinalizeParams fp;
  for (int k = 0; k < kNumLevels; ++k) {
    fp.lv[k].partials = pb.p[k];
    fp.lv[k].piv = pb.p[k] + (size_t)pairs * pb.chunks[k] * 5 * level_channels(k);
    fp.lv[k].chunks = k < levels_valid ? pb.chunks[k] : 0;
    fp.lv[k].C = level_channels(k);
    fp.lv[k].choff = level_offset(k);
    fp.lv[k].inv_n = 1.0 / ((double)h[k] * w[k]);
  }
  ProfScope ps(ctx, NQA_K_FINALIZE, st);
  ctx->launches += launch_dists_finalize(fp, ctx->alpha, ctx->beta, ctx->ab_normalize, pairs, score, s1s2, moments, pb.terms, st);
  return NQA_OK;
}

// levels 1..4 of the non-split trunks: L2pooling and the level's statistics in one pass (pool_stats_kernel)
// (bf16x3 keeps the separate pooling and moment kernels: a hi + lo version of the register-fed fused kernel was measured
// slower, 60.4 vs 34.2 + 21.4 ms per 64 pairs)
bool use_fused_pool(const nqa_ctx* ctx, int prec) { return ctx->fuse_pool_stats && prec != NQA_PREC_BF16X3; }

template <typename T>
size_t dists_mb_bytes(const nqa_ctx* ctx, int pairs, int H, int W, int prec) {
  return trunk_bytes<T>(2 * pairs, H, W, prec) + partial_bytes(pairs, H, W, prec, nullptr, false, use_fused_pool(ctx, prec));
}

template <typename T>
int dists_micro_batch(nqa_ctx* ctx, const float* x, const float* y, int pairs, int H, int W, int prec, float* score,
                      float* s1s2, cudaStream_t st) {
  ctx->arena.reset();
  TrunkBufs<T> bufs;
  trunk_alloc<T>(ctx, &bufs, 2 * pairs, H, W, prec);
  const bool fused_pool = use_fused_pool(ctx, prec);
  PartialBufs pb;
  partial_bytes(pairs, H, W, prec, &pb, false, fused_pool);
  partial_alloc(ctx, pairs, &pb);

  // level 0: the RAW frames (DISTS_pt.py:103 returns x itself, not the normalised tensor)
  {
    ProfScope ps(ctx, NQA_K_STATS, st);
    ctx->launches += launch_stats_planar(x, y, pairs, 3, H * W, pb.chunks[0], pb.p[0], st);
  }
  const int rc = run_trunk<T>(ctx, x, y, pairs, 2 * pairs, H, W, prec, bufs, st,
                              [&](int k, const T* f, const T* f_lo, int h, int w, T* pooled_dst) {
                                const int c = level_channels(k);
                                const size_t half = (size_t)pairs * h * w * c;
                                if (fused_pool && pooled_dst) {
                                  const size_t half_out = (size_t)pairs * pooled(h) * pooled(w) * c;
                                  ProfScope ps(ctx, NQA_K_POOL_STATS, st);
                                  ctx->launches += launch_pool_stats<T>(f, f + half, pooled_dst, pooled_dst + half_out, pairs,
                                                                        h, w, c, pb.chunks[k], pb.p[k], st, ctx->pool_stats_tma);
                                  return true;
                                }
                                ProfScope ps(ctx, NQA_K_STATS, st);
                                ctx->launches += launch_stats_nhwc<T>(f, f + half, f_lo, f_lo ? f_lo + half : nullptr, pairs,
                                                                     h * w, c, pb.chunks[k], pb.p[k], st);
                                return false;
                              });
  if (rc != NQA_OK) return rc;
  return finalize(ctx, pb, pairs, H, W, score, s1s2, st);
}

int dists_forward_device(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int prec, float* score,
                         float* s1s2, cudaStream_t st) {
  const int mb = pick_micro_pairs(ctx, B, H, W);
  const size_t need = align256(prec == NQA_PREC_FP32 ? dists_mb_bytes<float>(ctx, mb, H, W, prec)
                                                     : dists_mb_bytes<bf16>(ctx, mb, H, W, prec));
  const int n_mb = ceil_div(B, mb);
  const int lanes = (ctx->lanes > 1 && n_mb > 1) ? 2 : 1;
  int rc = ensure_arena(ctx, lanes * need);
  if (rc != NQA_OK) return rc;
  const size_t frame = (size_t)3 * H * W;
  if (lanes > 1) {   // fork: both lanes start after everything already enqueued on the caller's stream
    CUDA_TRY(cudaEventRecord(ctx->ev_fork, st));
    for (int l = 0; l < 2; ++l) CUDA_TRY(cudaStreamWaitEvent(ctx->lane_stream[l], ctx->ev_fork, 0));
  }
  const Arena whole = ctx->arena;
  int it = 0;
  for (int b0 = 0; b0 < B; b0 += mb, ++it) {
    const int pairs = B - b0 < mb ? B - b0 : mb;
    float* s12 = s1s2 ? s1s2 + (size_t)b0 * 2 * kTotalChannels : nullptr;
    const int lane = lanes > 1 ? (it & 1) : 0;
    cudaStream_t ls = lanes > 1 ? ctx->lane_stream[lane] : st;
    // each lane owns one half of the arena (dists_micro_batch resets and carves ctx->arena)
    ctx->arena.base = whole.base + (size_t)lane * need;
    ctx->arena.cap = need;
    ctx->arena.used = 0;
    rc = prec == NQA_PREC_FP32
             ? dists_micro_batch<float>(ctx, x + b0 * frame, y + b0 * frame, pairs, H, W, prec, score + b0, s12, ls)
             : dists_micro_batch<bf16>(ctx, x + b0 * frame, y + b0 * frame, pairs, H, W, prec, score + b0, s12, ls);
    ctx->arena = whole;
    if (rc != NQA_OK) return rc;
  }
  if (lanes > 1) {   // join
    for (int l = 0; l < 2; ++l) {
      CUDA_TRY(cudaEventRecord(ctx->ev_join[l], ctx->lane_stream[l]));
      CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_join[l], 0));
    }
  }
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

// ---- DISTS backward: one micro-batch ------------------------------------------------------------
// Recomputes the trunk forward with every activation kept (activation recomputation: nothing survives between the
// forward and the backward C-ABI calls), then walks the layers back.  See backward.cu for the math.
// gin_lo / gout_lo / mask_lo: bf16x3 carriers (all null for the single-carrier trunks)
template <typename T>
int conv_layer_bwd(nqa_ctx* ctx, int li, const T* gin, T* gout, int n, int h, int w, int epi, const T* mask, cudaStream_t st,
                   const T* gin_lo = nullptr, T* gout_lo = nullptr, const T* mask_lo = nullptr);
template <>
int conv_layer_bwd<float>(nqa_ctx* ctx, int li, const float* gin, float* gout, int n, int h, int w, int epi,
                          const float* mask, cudaStream_t st, const float*, float*, const float*) {
  ProfScope ps(ctx, NQA_K_CONV_F32, st);
  ctx->launches += launch_conv3x3_f32(gin, gout, ctx->w_bwd_f32[li], nullptr, n, h, w, kConvCout[li], kConvCin[li], st, epi, mask);
  return NQA_OK;
}
template <>
int conv_layer_bwd<bf16>(nqa_ctx* ctx, int li, const bf16* gin, bf16* gout, int n, int h, int w, int epi, const bf16* mask,
                         cudaStream_t st, const bf16* gin_lo, bf16* gout_lo, const bf16* mask_lo) {
  ConvTcPlan plan;
  const char* err = "";
  const bool split = gin_lo != nullptr;   // bf16x3: hi/lo gradients and weights, three MMAs per k-step (per-tap kernel)
  // roles of the channel counts are swapped; the bias pointer is only staged (512 floats), never applied
  if (cached_plan(ctx, &plan, gin, gin_lo, ctx->w_bwd_hi[li], split ? ctx->w_bwd_lo[li] : nullptr, ctx->bias[kNumConvs - 1],
                  gout, gout_lo, n, h, w, kConvCout[li], kConvCin[li], split, &err))
    return fail(NQA_ERR_CUDA, "%s (backward of layer %d, %dx%dx%d)", err, li, n, h, w);
  plan.epi = epi;
  plan.mask = mask;
  plan.mask_lo = mask_lo;
  ProfScope ps(ctx, NQA_K_CONV_TC, st);
  const int nl = launch_conv3x3_tc(plan, st);
  if (nl <= 0) return fail(NQA_ERR_CUDA, "backward convolution of layer %d could not be launched", li);
  ctx->launches += nl;
  return NQA_OK;
}

// Backward of conv1_1 + input normalisation: gz = gradient w.r.t. the pre-ReLU conv1_1 output, NHWC [n,H,W,64];
// writes the frame gradients (planes).  bf16 trunk: the halo convolution kernel with the three input channels as
// output rows and the kEpiFrameGrad epilogue (the CUDA-core kernel is ~2.5x slower); fp32 trunk: CUDA cores.
template <typename T>
int conv1_1_bwd(nqa_ctx* ctx, const T* gz, T* scratch, int pairs, int n, int H, int W, float* gx, float* gy, cudaStream_t st,
                const T* gz_lo = nullptr);
template <>
int conv1_1_bwd<float>(nqa_ctx* ctx, const float* gz, float*, int pairs, int n, int H, int W, float* gx, float* gy,
                       cudaStream_t st, const float*) {
  ctx->launches += launch_conv1_1_bwd<float>(gz, ctx->w11_bwd_host, pairs, n, H, W, gx, gy, st);
  return NQA_OK;
}
template <>
int conv1_1_bwd<bf16>(nqa_ctx* ctx, const bf16* gz, bf16* scratch, int pairs, int n, int H, int W, float* gx, float* gy,
                      cudaStream_t st, const bf16* gz_lo) {
  ConvTcPlan plan;
  const char* err = "";
  if (gz_lo) {   // bf16x3: the CUDA-core kernel sums the two carriers in fp32
    ProfScope psb(ctx, NQA_K_BACKWARD, st);
    ctx->launches += launch_conv1_1_bwd<bf16>(gz, ctx->w11_bwd_host, pairs, n, H, W, gx, gy, st, gz_lo);
    return NQA_OK;
  }
  if (ctx->conv11_bwd_tc &&
      cached_plan(ctx, &plan, gz, nullptr, ctx->w11_bwd_hi, nullptr, ctx->bias[kNumConvs - 1], scratch, nullptr, n, H, W, 64, 64,
                  false, &err) == 0 &&
      plan.halo && !plan.pair) {
    plan.epi = kEpiFrameGrad;
    plan.fg_x = gx; plan.fg_y = gy; plan.fg_nx = pairs;
    ProfScope ps(ctx, NQA_K_CONV_TC, st);
    const int nl = launch_conv3x3_tc(plan, st);
    if (nl <= 0) return fail(NQA_ERR_CUDA, "backward convolution of conv1_1 could not be launched");
    ctx->launches += nl;
    return NQA_OK;
  }
  ctx->launches += launch_conv1_1_bwd<bf16>(gz, ctx->w11_bwd_host, pairs, n, H, W, gx, gy, st);
  return NQA_OK;
}


template <typename T>
size_t saved_bytes(int pairs, int H, int W, int prec = NQA_PREC_BF16) {
  int h[kNumLevels], w[kNumLevels];
  level_dims(H, W, h, w);
  const size_t n = 2 * (size_t)pairs;
  size_t total = 0;
  int li = 0;
  for (int stage = 0; stage < 5; ++stage) {
    const size_t hw = (size_t)h[stage + 1] * w[stage + 1];
    for (int j = 0; j < kStageConvs[stage]; ++j, ++li) total += align256(n * hw * kConvCout[li] * sizeof(T));
    if (stage < 4) total += align256(n * (size_t)h[stage + 2] * w[stage + 2] * kConvCout[li - 1] * sizeof(T));
  }
  return total * carriers(prec) + align256((size_t)pairs * kTotalChannels * 5 * sizeof(float)) + 4096;
}

// workspace of one micro-batch besides the saved activations: statistics partials, coefficient table, gradient ping-pong
template <typename T>
size_t bwd_work_bytes(int pairs, int H, int W, int prec) {
  const size_t n = 2 * (size_t)pairs;
  return 2 * carriers(prec) * align256(n * max_act_elems(H, W) * sizeof(T)) +
         partial_bytes(pairs, H, W, prec, nullptr, false, prec != NQA_PREC_BF16X3) +
         align256((size_t)pairs * 5 * kCoefStride * sizeof(float)) + align256(pairs * sizeof(float)) + 8192;
}

template <typename T>
size_t dists_bwd_bytes(const nqa_ctx*, int pairs, int H, int W, int prec) {
  return saved_bytes<T>(pairs, H, W, prec) + bwd_work_bytes<T>(pairs, H, W, prec);
}

// Trunk forward of one micro-batch with every activation kept in `store` (not the fused stage-1 kernel: relu1_1 is
// needed by the backward), statistics and scores as in dists_micro_batch.  Statistics partials come from ctx->arena
// (the caller resets it).
template <typename T>
int trunk_forward_keep(nqa_ctx* ctx, Arena& store, const float* x, const float* y, int pairs, int H, int W, int prec,
                       bool through_trunk, SavedActs* sv, float* score, float* s1s2, cudaStream_t st) {
  int lh[kNumLevels], lw[kNumLevels];
  level_dims(H, W, lh, lw);
  const int n = 2 * pairs;
  const bool split = prec == NQA_PREC_BF16X3;   // bf16 hi + residual lo carriers for every kept tensor
  PartialBufs pb;
  partial_bytes(pairs, H, W, prec, &pb, false, !split);
  partial_alloc(ctx, pairs, &pb);
  sv->pairs = pairs;
  sv->moments = static_cast<float*>(store.take((size_t)pairs * kTotalChannels * 5 * sizeof(float)));
  {  // level 0: the RAW frames
    ProfScope ps(ctx, NQA_K_STATS, st);
    ctx->launches += launch_stats_planar(x, y, pairs, 3, H * W, pb.chunks[0], pb.p[0], st);
  }
  if (through_trunk) {
    int li = 0;
    const T* cur = nullptr;
    const T* cur_lo = nullptr;
    for (int stage = 0; stage < 5; ++stage) {
      const int h = lh[stage + 1], w = lw[stage + 1];
      for (int j = 0; j < kStageConvs[stage]; ++j, ++li) {
        const size_t bytes = (size_t)n * h * w * kConvCout[li] * sizeof(T);
        T* out = static_cast<T*>(store.take(bytes));
        T* out_lo = split ? static_cast<T*>(store.take(bytes)) : nullptr;
        sv->act[li] = out;
        sv->act_lo[li] = out_lo;
        if (li == 0) {
          conv1_1_layer<T>(ctx, x, y, pairs, n, h, w, out, out_lo, st);
        } else {
          const int rc = conv_layer<T>(ctx, li, cur, cur_lo, out, out_lo, n, h, w, prec, st);
          if (rc != NQA_OK) return rc;
        }
        cur = out;
        cur_lo = out_lo;
      }
      const int k = stage + 1, c = level_channels(k);
      const size_t half = (size_t)pairs * h * w * c;
      if (split) {
        // bf16x3: statistics and pooling as separate kernels on the hi + lo carriers
        {
          ProfScope ps(ctx, NQA_K_STATS, st);
          ctx->launches += launch_stats_nhwc<T>(cur, cur + half, cur_lo, cur_lo + half, pairs, h * w, c, pb.chunks[k], pb.p[k], st);
        }
        if (stage < 4) {
          const size_t pbytes = (size_t)n * lh[k + 1] * lw[k + 1] * c * sizeof(T);
          T* pooled_t = static_cast<T*>(store.take(pbytes));
          T* pooled_lo = static_cast<T*>(store.take(pbytes));
          sv->pool[stage] = pooled_t;
          sv->pool_lo[stage] = pooled_lo;
          ProfScope ps(ctx, NQA_K_L2POOL, st);
          const int nl = launch_l2pool_nhwc<T>(cur, cur_lo, pooled_t, pooled_lo, n, h, w, c, st);
          if (nl <= 0) return fail(NQA_ERR_SHAPE, "L2pooling: %d images of %dx%d exceed the launch grid", n, h, w);
          ctx->launches += nl;
          cur = pooled_t;
          cur_lo = pooled_lo;
        }
      } else if (stage < 4) {
        const size_t half_out = (size_t)pairs * lh[k + 1] * lw[k + 1] * c;
        T* pooled_t = static_cast<T*>(store.take(2 * half_out * sizeof(T)));
        sv->pool[stage] = pooled_t;
        ProfScope ps(ctx, NQA_K_POOL_STATS, st);
        ctx->launches += launch_pool_stats<T>(cur, cur + half, pooled_t, pooled_t + half_out, pairs, h, w, c, pb.chunks[k], pb.p[k],
                                              st, ctx->pool_stats_tma);
        cur = pooled_t;
      } else {
        ProfScope ps(ctx, NQA_K_STATS, st);
        ctx->launches += launch_stats_nhwc<T>(cur, cur + half, nullptr, nullptr, pairs, h * w, c, pb.chunks[k], pb.p[k], st);
      }
    }
  }
  const int levels_valid = through_trunk ? kNumLevels : 1;
  return finalize(ctx, pb, pairs, H, W, score, s1s2, st, sv->moments, levels_valid);
}

// Walks the layers back from the saved activations of one micro-batch.  See backward.cu for the math.
template <typename T>
int trunk_backward_walk(nqa_ctx* ctx, const float* x, const float* y, int pairs, int H, int W, const SavedActs& sv,
                        const float* grad_score, bool through_trunk, float* gx, float* gy, cudaStream_t st) {
  int lh[kNumLevels], lw[kNumLevels];
  level_dims(H, W, lh, lw);
  const int n = 2 * pairs;
  T* const* act = reinterpret_cast<T* const*>(sv.act);
  T* const* pool = reinterpret_cast<T* const*>(sv.pool);
  float* coef = static_cast<float*>(ctx->arena.take((size_t)pairs * 5 * kCoefStride * sizeof(float)));
  GradLevels gl;
  for (int k = 0; k < kNumLevels; ++k) gl.inv_n[k] = 1.0 / ((double)lh[k] * lw[k]);
  const int levels_valid = through_trunk ? kNumLevels : 1;
  int rc = NQA_OK;
  {
    ProfScope psb(ctx, NQA_K_BACKWARD, st);
    ctx->launches += launch_stat_coef(sv.moments, ctx->alpha, ctx->beta, ctx->ab_normalize, grad_score, gl, levels_valid, pairs, coef, st);
  }

  if (through_trunk) {
    const size_t gsz = (size_t)n * max_act_elems(H, W) * sizeof(T);
    T* const* act_lo = reinterpret_cast<T* const*>(sv.act_lo);
    T* const* pool_lo = reinterpret_cast<T* const*>(sv.pool_lo);
    const bool split = sv.act_lo[0] != nullptr;   // bf16x3 tape: every tensor travels as hi + lo
    T* g[2] = {static_cast<T*>(ctx->arena.take(gsz)), static_cast<T*>(ctx->arena.take(gsz))};
    T* gl[2] = {nullptr, nullptr};
    if (split) { gl[0] = static_cast<T*>(ctx->arena.take(gsz)); gl[1] = static_cast<T*>(ctx->arena.take(gsz)); }
    int cur = 0;
    {  // last level: statistics gradient only
      const int h = lh[5], w = lw[5], c = level_channels(5);
      const size_t half = (size_t)pairs * h * w * c;
      LevelGradLo<T> lo;
      if (split) { lo.fx = act_lo[12]; lo.fy = act_lo[12] + half; lo.gx = gl[cur]; lo.gy = gl[cur] + half; }
      ProfScope psb(ctx, NQA_K_BACKWARD, st);
      const int nl = launch_level_grad<T>(act[12], act[12] + half, nullptr, nullptr, coef, 5, pairs, h, w, c, g[cur],
                                          g[cur] + half, st, split ? &lo : nullptr);
      if (nl <= 0) return fail(NQA_ERR_INVALID, "level gradient: unsupported shape %dx%dx%d", h, w, c);
      ctx->launches += nl;
    }
    int li = kNumConvs - 1;
    for (int stage = 4; stage >= 0; --stage) {
      const int h = lh[stage + 1], w = lw[stage + 1];
      for (int j = kStageConvs[stage] - 1; j >= 0; --j, --li) {
        if (li == 0) break;
        // first conv of a stage: its input is the pooled output of the stage below; the gradient w.r.t. that tensor
        // is handed on as grad / pooled (the only form the pooling backward needs)
        const bool first = j == 0;
        rc = conv_layer_bwd<T>(ctx, li, g[cur], g[cur ^ 1], n, h, w, first ? kEpiDivide : kEpiMask,
                               first ? pool[stage - 1] : act[li - 1], st, split ? gl[cur] : nullptr,
                               split ? gl[cur ^ 1] : nullptr, split && first ? pool_lo[stage - 1] : nullptr);
        if (rc != NQA_OK) return rc;
        cur ^= 1;
      }
      if (stage == 0) break;
      // g[cur] = gradient w.r.t. the pooled input of this stage -> gradient w.r.t. the pre-ReLU output of the last
      // conv of the stage below (pooling backward + that level's statistics term + ReLU mask)
      const int k = stage, hp = lh[k], wp = lw[k], c = level_channels(k);
      const size_t half = (size_t)pairs * hp * wp * c, half_pool = (size_t)pairs * h * w * c;
      const T* feat = act[li];                            // li now indexes the last conv of the stage below
      LevelGradLo<T> lo;
      if (split) {
        lo.fx = act_lo[li]; lo.fy = act_lo[li] + half;
        lo.qpoolx = gl[cur]; lo.qpooly = gl[cur] + half_pool;
        lo.gx = gl[cur ^ 1]; lo.gy = gl[cur ^ 1] + half;
      }
      ProfScope psb(ctx, NQA_K_BACKWARD, st);
      const int nl = launch_level_grad<T>(feat, feat + half, g[cur], g[cur] + half_pool, coef, k, pairs, hp, wp, c, g[cur ^ 1],
                                          g[cur ^ 1] + half, st, split ? &lo : nullptr);
      if (nl <= 0) return fail(NQA_ERR_INVALID, "level gradient: unsupported shape %dx%dx%d", hp, wp, c);
      ctx->launches += nl;
      cur ^= 1;
    }
    rc = conv1_1_bwd<T>(ctx, g[cur], g[cur ^ 1], pairs, n, H, W, gx, gy, st, split ? gl[cur] : nullptr);
    if (rc != NQA_OK) return rc;
  }
  ProfScope psb(ctx, NQA_K_BACKWARD, st);
  ctx->launches += launch_stat_grad_planar(x, y, coef, 0, pairs, 3, H * W, gx, gy, through_trunk, st);   // level 0
  return NQA_OK;
}

// ---- DISTS backward by activation recomputation: one micro-batch ------------------------------------------------
// Recomputes the trunk forward with every activation kept in the arena (nothing survives between nqa_dists_forward and
// nqa_dists_backward), then walks the layers back.
template <typename T>
int dists_backward_micro_batch(nqa_ctx* ctx, const float* x, const float* y, int pairs, int H, int W, int prec,
                               const float* grad_score, bool through_trunk, float* gx, float* gy, cudaStream_t st) {
  ctx->arena.reset();
  float* score_scratch = static_cast<float*>(ctx->arena.take(pairs * sizeof(float)));
  SavedActs sv;
  int rc = trunk_forward_keep<T>(ctx, ctx->arena, x, y, pairs, H, W, prec, through_trunk, &sv, score_scratch, nullptr, st);
  if (rc != NQA_OK) return rc;
  return trunk_backward_walk<T>(ctx, x, y, pairs, H, W, sv, grad_score, through_trunk, gx, gy, st);
}

// ---- tape: activations kept from nqa_dists_forward_saved to nqa_dists_backward_saved -----------------------------
Tape* find_tape(nqa_ctx* ctx, unsigned long long id) {
  for (auto& t : ctx->tapes)
    if (t.id == id) return &t;
  return nullptr;
}

void release_tape(nqa_ctx* ctx, unsigned long long id) {
  for (size_t i = 0; i < ctx->tapes.size(); ++i)
    if (ctx->tapes[i].id == id) {
      ctx->tape_bufs[ctx->tapes[i].buf].in_use = false;
      ctx->tapes.erase(ctx->tapes.begin() + i);
      return;
    }
}

// Index of a free tape buffer of at least `bytes` (smallest that fits, else a new allocation); -1 if the tape limit
// would be exceeded or the allocation fails (the caller then falls back to activation recomputation).
int acquire_tape_buf(nqa_ctx* ctx, size_t bytes) {
  int best = -1;
  size_t held = 0;
  for (size_t i = 0; i < ctx->tape_bufs.size(); ++i) {
    const TapeBuf& b = ctx->tape_bufs[i];
    held += b.cap;
    if (!b.in_use && b.cap >= bytes && (best < 0 || b.cap < ctx->tape_bufs[best].cap)) best = (int)i;
  }
  if (best >= 0) {
    ctx->tape_bufs[best].in_use = true;
    return best;
  }
  // drop idle buffers that are too small before growing
  for (size_t i = 0; i < ctx->tape_bufs.size(); ++i) {
    TapeBuf& b = ctx->tape_bufs[i];
    if (!b.in_use && b.base) {
      cudaDeviceSynchronize();
      cudaFree(b.base);
      held -= b.cap;
      b.base = nullptr;
      b.cap = 0;
    }
  }
  if (held + bytes > ctx->tape_limit) return -1;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  for (size_t i = 0; i < ctx->tape_bufs.size(); ++i)
    if (!ctx->tape_bufs[i].base) {
      ctx->tape_bufs[i] = TapeBuf{static_cast<char*>(p), bytes, true};
      return (int)i;
    }
  ctx->tape_bufs.push_back(TapeBuf{static_cast<char*>(p), bytes, true});
  return (int)ctx->tape_bufs.size() - 1;
}

// Host-buffer entry points: frames are streamed host -> device in micro-batches on the copy stream, double-buffered
// against the compute stream; `score(dx, dy, pairs, d_out, stream)` enqueues the scoring of one staged micro-batch and
// writes `pairs` scores to each of d_out[0..n_out).  Returns when the n_out score vectors are in out_host[0..n_out).
template <typename F>
int host_stream_run(nqa_ctx* ctx, const float* x_host, const float* y_host, int B, int H, int W, int mb, int n_out,
                    float* const* out_host, F&& score) {
  const size_t frame = (size_t)3 * H * W;
  const size_t slot_floats = 2 * (size_t)mb * frame + 2 * (size_t)B;  // x frames, y frames (+ the scores in slot 0)
  if (slot_floats > ctx->stage_cap) {
    CUDA_TRY(cudaDeviceSynchronize());
    for (int i = 0; i < 2; ++i) {
      if (ctx->stage_in[i]) CUDA_TRY(cudaFree(ctx->stage_in[i]));
      ctx->stage_in[i] = nullptr;
    }
    ctx->stage_cap = 0;
    for (int i = 0; i < 2; ++i)
      if (cudaMalloc(&ctx->stage_in[i], slot_floats * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return fail(NQA_ERR_NOMEM, "staging allocation of %zu bytes failed", slot_floats * sizeof(float));
      }
    ctx->stage_cap = slot_floats;
  }
  float* d_scores = ctx->stage_in[0] + 2 * (size_t)mb * frame;
  // the arena, the staging slots and the channel weights are shared with the device entry points: order this call
  // after whatever the caller enqueued earlier on its own stream (and later calls after this one)
  StreamOrder so(ctx, ctx->compute_stream);
  if (so.rc != NQA_OK) return so.rc;
  CUDA_TRY(cudaEventRecord(ctx->ev_fork, ctx->compute_stream));
  CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
  // The first chunk is small so that the copy/compute pipeline fills quickly (its H2D copy is the only one that is
  // not hidden behind compute); scores do not depend on the chunking (batch-independent reductions).
  int rc = NQA_OK, it = 0;
  for (int b0 = 0; b0 < B; ++it) {
    const int slot = it & 1;
    const int want = it == 0 ? (mb > 2 ? 2 : mb) : mb;
    const int pairs = B - b0 < want ? B - b0 : want;
    float* dx = ctx->stage_in[slot];
    float* dy = dx + (size_t)mb * frame;
    if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[slot], 0));
    CUDA_TRY(cudaMemcpyAsync(dx, x_host + b0 * frame, pairs * frame * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(cudaMemcpyAsync(dy, y_host + b0 * frame, pairs * frame * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(cudaEventRecord(ctx->ev_copied[slot], ctx->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->compute_stream, ctx->ev_copied[slot], 0));
    float* d_out[2] = {d_scores + b0, d_scores + B + b0};
    if ((rc = score(dx, dy, pairs, d_out, ctx->compute_stream)) != NQA_OK) break;
    CUDA_TRY(cudaEventRecord(ctx->ev_consumed[slot], ctx->compute_stream));
    b0 += pairs;
  }
  if (rc == NQA_OK)
    for (int o = 0; o < n_out; ++o)
      CUDA_TRY(cudaMemcpyAsync(out_host[o], d_scores + (size_t)o * B, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost,
                               ctx->compute_stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->compute_stream));
  if (rc != NQA_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int check_ready(nqa_ctx* ctx, bool need_ab) {
  if (!ctx) return fail(NQA_ERR_INVALID, "null context");
  if (!ctx->have_weights) return fail(NQA_ERR_STATE, "VGG weights not set (nqa_set_vgg_weights)");
  if (need_ab && !ctx->have_alpha_beta) return fail(NQA_ERR_STATE, "alpha/beta not set (nqa_set_alpha_beta)");
  return NQA_OK;
}

int check_prec(int prec) {
  if (prec != NQA_PREC_FP32 && prec != NQA_PREC_BF16 && prec != NQA_PREC_BF16X3)
    return fail(NQA_ERR_INVALID, "unknown precision %d", prec);
  return NQA_OK;
}

}  // namespace

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" {

const char* nqa_last_error(void) { return g_last_error.c_str(); }
int nqa_version(void) { return 100; }

int nqa_ctx_create(int device, nqa_ctx** out) {
  if (!out) return fail(NQA_ERR_INVALID, "null out pointer");
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
    cudaGetLastError();
    return fail(NQA_ERR_CUDA, "no CUDA device available (nerfqa_b200 has no CPU fallback)");
  }
  if (device < 0 || device >= count) return fail(NQA_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(NQA_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(NQA_ERR_CUDA, "device %d is sm_%d%d; nerfqa_b200 is built for sm_100a (B200) only", device, prop.major,
                prop.minor);
  nqa_ctx* ctx = new nqa_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->conv_cfg = conv_tc_config_from_env(ctx->sm_count);
  if (const char* e = getenv("NQA_POOL_STATS_TMA")) ctx->pool_stats_tma = atoi(e) != 0;
  if (const char* e = getenv("NQA_ADISTS_FUSED_SUMS")) ctx->adists_fused_sums = atoi(e) != 0;
  if (const char* e = getenv("NQA_ADISTS_STRIPS")) ctx->adists_strips = atoi(e);
  if (const char* e = getenv("NQA_ADISTS_MMA")) ctx->adists_opt.use_mma = atoi(e) != 0;
  if (const char* e = getenv("NQA_ADISTS_MMA_TERMS")) ctx->adists_opt.mma_terms = atoi(e);
  if (const char* e = getenv("NQA_ADISTS_ONE_TERM_MIN_PIXELS")) ctx->adists_opt.one_term_min_pixels = atoll(e);
  CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_last, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_graph_in, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_graph_out, cudaEventDisableTiming));
  if (const char* e = getenv("NQA_GRAPHS")) ctx->use_graphs = atoi(e) != 0;
  if (const char* e = getenv("NQA_GRAPH_MAX_PIXELS")) ctx->graph_max_pixels = atoll(e);
  if (const char* e = getenv("NQA_STAGE1_FUSED")) ctx->fuse_stage1 = atoi(e) != 0;
  if (const char* e = getenv("NQA_POOL_STATS_FUSED")) ctx->fuse_pool_stats = atoi(e) != 0;
  if (const char* e = getenv("NQA_STAGE1_PAIR")) ctx->stage1_pair = atoi(e) != 0;
  if (const char* e = getenv("NQA_CONV11_BWD_TC")) ctx->conv11_bwd_tc = atoi(e) != 0;
  if (const char* e = getenv("NQA_TAPE_MAX_MB")) ctx->tape_limit = (size_t)atoll(e) << 20;
  if (const char* e = getenv("NQA_STAGE1_A2_LINEAR")) ctx->stage1_a2_linear = atoi(e) != 0;
  if (const char* e = getenv("NQA_MICRO_IMAGES")) ctx->micro_images = atoi(e) > 0 ? atoi(e) : 16;
  if (const char* e = getenv("NQA_LANES")) ctx->lanes = atoi(e) == 2 ? 2 : 1;
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->lane_stream[i], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->compute_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaMalloc(&ctx->alpha, kTotalChannels * sizeof(float)));
  CUDA_TRY(cudaMalloc(&ctx->beta, kTotalChannels * sizeof(float)));
  CUDA_TRY(cudaMalloc(&ctx->w11, 27 * 64 * sizeof(float)));
  CUDA_TRY(cudaMalloc(&ctx->w11_img, 2 * 64 * 32 * sizeof(bf16)));
  CUDA_TRY(cudaMalloc(&ctx->w11_bwd, 27 * 64 * sizeof(float)));
  CUDA_TRY(cudaMalloc(&ctx->w11_bwd_hi, 64 * 9 * 64 * sizeof(bf16)));
  CUDA_TRY(cudaMemset(ctx->w11_bwd_hi, 0, 64 * 9 * 64 * sizeof(bf16)));
  for (int i = 0; i < kNumConvs; ++i) {
    const size_t nw = (size_t)kConvCout[i] * kConvCin[i] * 9;
    CUDA_TRY(cudaMalloc(&ctx->bias[i], kConvCout[i] * sizeof(float)));
    if (i > 0) {
      CUDA_TRY(cudaMalloc(&ctx->w_f32[i], nw * sizeof(float)));
      CUDA_TRY(cudaMalloc(&ctx->w_hi[i], nw * sizeof(bf16)));
      CUDA_TRY(cudaMalloc(&ctx->w_lo[i], nw * sizeof(bf16)));
      CUDA_TRY(cudaMalloc(&ctx->w_bwd_f32[i], nw * sizeof(float)));
      CUDA_TRY(cudaMalloc(&ctx->w_bwd_hi[i], nw * sizeof(bf16)));
      CUDA_TRY(cudaMalloc(&ctx->w_bwd_lo[i], nw * sizeof(bf16)));
    }
  }
  *out = ctx;
  return NQA_OK;
}

int nqa_ctx_destroy(nqa_ctx* ctx) {
  if (!ctx) return NQA_OK;
  DeviceGuard guard(ctx->device);
  cudaDeviceSynchronize();
  cudaFree(ctx->alpha);
  cudaFree(ctx->beta);
  cudaFree(ctx->w11);
  cudaFree(ctx->w11_img);
  cudaFree(ctx->w11_bwd);
  cudaFree(ctx->w11_bwd_hi);
  for (auto& b : ctx->tape_bufs) cudaFree(b.base);
  for (int i = 0; i < kNumConvs; ++i) {
    cudaFree(ctx->w_bwd_f32[i]);
    cudaFree(ctx->w_bwd_hi[i]);
    cudaFree(ctx->w_bwd_lo[i]);
    cudaFree(ctx->bias[i]);
    cudaFree(ctx->w_f32[i]);
    cudaFree(ctx->w_hi[i]);
    cudaFree(ctx->w_lo[i]);
  }
  cudaFree(ctx->arena.base);
  for (int i = 0; i < 2; ++i) {
    cudaFree(ctx->stage_in[i]);
    if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    if (ctx->ev_consumed[i]) cudaEventDestroy(ctx->ev_consumed[i]);
  }
  for (auto& r : ctx->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : ctx->prof_pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (ctx->lane_stream[i]) cudaStreamDestroy(ctx->lane_stream[i]);
    if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_last) cudaEventDestroy(ctx->ev_last);
  if (ctx->ev_graph_in) cudaEventDestroy(ctx->ev_graph_in);
  if (ctx->ev_graph_out) cudaEventDestroy(ctx->ev_graph_out);
  drop_graphs(ctx);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->compute_stream) cudaStreamDestroy(ctx->compute_stream);
  cudaGetLastError();
  delete ctx;
  return NQA_OK;
}

int nqa_set_vgg_weights(nqa_ctx* ctx, const float* const w[NQA_NUM_CONVS], const float* const b[NQA_NUM_CONVS]) {
  if (!ctx || !w || !b) return fail(NQA_ERR_INVALID, "null argument");
  for (int i = 0; i < kNumConvs; ++i)
    if (!w[i] || !b[i]) return fail(NQA_ERR_INVALID, "null weight/bias pointer for conv %d", i);
  DeviceGuard guard(ctx->device);
  float* tmp = nullptr;
  const size_t max_w = (size_t)512 * 512 * 9;
  CUDA_TRY(cudaMalloc(&tmp, max_w * sizeof(float)));
  for (int i = 0; i < kNumConvs; ++i) {
    const size_t nw = (size_t)kConvCout[i] * kConvCin[i] * 9;
    cudaError_t e = cudaMemcpy(tmp, w[i], nw * sizeof(float), cudaMemcpyDefault);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->bias[i], b[i], kConvCout[i] * sizeof(float), cudaMemcpyDefault);
    if (e != cudaSuccess) {
      cudaFree(tmp);
      return fail(NQA_ERR_CUDA, "copying conv %d parameters failed: %s", i, cudaGetErrorString(e));
    }
    if (i == 0) {
      pack_w11_kernel<<<(27 * 64 + 255) / 256, 256>>>(tmp, ctx->w11);
      ctx->launches += launch_pack_w11_tc(tmp, ctx->w11_img, 0);
      ctx->launches += launch_pack_w11_bwd(tmp, ctx->w11_bwd, 0);
      // rows ci = 0..2 of the [64][9][64] tensor-pipe image (the other rows stay zero)
      ctx->launches += launch_pack_w_bwd(tmp, 3, 64, nullptr, ctx->w11_bwd_hi, 0);
    } else {
      pack_w_kernel<<<1024, 256>>>(tmp, kConvCin[i], kConvCout[i], ctx->w_f32[i], ctx->w_hi[i], ctx->w_lo[i]);
      ctx->launches += launch_pack_w_bwd(tmp, kConvCin[i], kConvCout[i], ctx->w_bwd_f32[i], ctx->w_bwd_hi[i], 0, ctx->w_bwd_lo[i]);
    }
    ctx->launches += 1;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      cudaFree(tmp);
      return fail(NQA_ERR_CUDA, "packing conv %d failed: %s", i, cudaGetErrorString(e));
    }
  }
  cudaFree(tmp);
  CUDA_TRY(cudaMemcpy(&ctx->w11_bwd_host, ctx->w11_bwd, sizeof(Conv11BwdWeights), cudaMemcpyDeviceToHost));
  if (conv_tc_make_weight_map(&ctx->w12_map_pair, ctx->w_hi[1], 64, 9 * 64, 32) != 0 ||
      conv_tc_make_weight_map(&ctx->w12_map, ctx->w_hi[1], 64, 9 * 64, 64) != 0)
    return fail(NQA_ERR_CUDA, "cuTensorMapEncodeTiled failed for the conv1_2 weights");
  ctx->have_weights = true;
  return NQA_OK;
}

int nqa_set_alpha_beta_async(nqa_ctx* ctx, const float* alpha, const float* beta, void* stream) {
  if (!ctx || !alpha || !beta) return fail(NQA_ERR_INVALID, "null argument");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ordered after earlier scoring calls that still read the old weights (on whatever stream they ran) and before
  // later ones: copies on the caller's stream, cross-stream hand-over through the context's ordering event
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  CUDA_TRY(cudaMemcpyAsync(ctx->alpha, alpha, kTotalChannels * sizeof(float), cudaMemcpyDefault, st));
  CUDA_TRY(cudaMemcpyAsync(ctx->beta, beta, kTotalChannels * sizeof(float), cudaMemcpyDefault, st));
  ctx->have_alpha_beta = true;
  ctx->ab_normalize = true;
  return NQA_OK;
}

int nqa_set_channel_weights_async(nqa_ctx* ctx, const float* a, const float* b, void* stream) {
  const int rc = nqa_set_alpha_beta_async(ctx, a, b, stream);
  if (rc == NQA_OK) ctx->ab_normalize = false;
  return rc;
}

// Synchronous forms: the copies run on the legacy default stream (ordered against the context's other streams like
// every call) and the call returns when they are done, so host buffers may be reused at once.
int nqa_set_alpha_beta(nqa_ctx* ctx, const float* alpha, const float* beta) {
  const int rc = nqa_set_alpha_beta_async(ctx, alpha, beta, nullptr);
  if (rc != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  return NQA_OK;
}

int nqa_set_channel_weights(nqa_ctx* ctx, const float* a, const float* b) {
  const int rc = nqa_set_alpha_beta(ctx, a, b);
  if (rc == NQA_OK) ctx->ab_normalize = false;
  return rc;
}

int nqa_dists_forward(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int precision, float* score_out,
                      float* s1s2_out, void* stream) {
  int rc = check_ready(ctx, true);
  if (rc != NQA_OK) return rc;
  if (!x || !y || !score_out) return fail(NQA_ERR_INVALID, "null tensor pointer");
  if ((rc = check_prec(precision)) != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  if (s1s2_out) return dists_forward_device(ctx, x, y, B, H, W, precision, score_out, s1s2_out, st);
  float* outs[1] = {score_out};
  return graphed_call(ctx, 0, x, y, B, H, W, precision, 0, outs, 1, st,
                      [&](const float* gx, const float* gy, float* const* go, cudaStream_t gs) {
                        return dists_forward_device(ctx, gx, gy, B, H, W, precision, go[0], nullptr, gs);
                      });
}

int nqa_dists_forward_host(nqa_ctx* ctx, const float* x_host, const float* y_host, int B, int H, int W, int precision,
                           float* score_host) {
  int rc = check_ready(ctx, true);
  if (rc != NQA_OK) return rc;
  if (!x_host || !y_host || !score_host) return fail(NQA_ERR_INVALID, "null tensor pointer");
  if ((rc = check_prec(precision)) != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  const int mb = pick_micro_pairs(ctx, B, H, W);
  const size_t need = precision == NQA_PREC_FP32 ? dists_mb_bytes<float>(ctx, mb, H, W, precision)
                                                 : dists_mb_bytes<bf16>(ctx, mb, H, W, precision);
  if ((rc = ensure_arena(ctx, need)) != NQA_OK) return rc;
  float* outs[1] = {score_host};
  return host_stream_run(ctx, x_host, y_host, B, H, W, mb, 1, outs,
                         [&](const float* dx, const float* dy, int pairs, float* const* d_out, cudaStream_t st) {
                           return precision == NQA_PREC_FP32
                                      ? dists_micro_batch<float>(ctx, dx, dy, pairs, H, W, precision, d_out[0], nullptr, st)
                                      : dists_micro_batch<bf16>(ctx, dx, dy, pairs, H, W, precision, d_out[0], nullptr, st);
                         });
}

}  // extern "C"

template <typename T>
static int features_impl(nqa_ctx* ctx, const float* x, int B, int H, int W, int prec, float* const* feats_out,
                         cudaStream_t st) {
  int mb = 2 * pick_micro_pairs(ctx, (B + 1) / 2, H, W);
  if (mb > B) mb = B;
  int rc = ensure_arena(ctx, trunk_bytes<T>(mb, H, W, prec));
  if (rc != NQA_OK) return rc;
  const size_t frame = (size_t)3 * H * W;
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int n = B - b0 < mb ? B - b0 : mb;
    ctx->arena.reset();
    TrunkBufs<T> bufs;
    trunk_alloc<T>(ctx, &bufs, n, H, W, prec);
    rc = run_trunk<T>(ctx, x + b0 * frame, x, n, n, H, W, prec, bufs, st,
                      [&](int k, const T* f, const T* f_lo, int h, int w, T*) {
                        float* dst = feats_out[k - 1] + (size_t)b0 * level_channels(k) * h * w;
                        level_to_nchw<T>(ctx, f, f_lo, dst, n, level_channels(k), h, w, st);
                        return false;
                      });
    if (rc != NQA_OK) return rc;
  }
  return NQA_OK;
}

extern "C" {

int nqa_dists_features(nqa_ctx* ctx, const float* x, int B, int H, int W, int precision,
                       float* const feats_out[NQA_NUM_LEVELS - 1], void* stream) {
  int rc = check_ready(ctx, false);
  if (rc != NQA_OK) return rc;
  if (!x || !feats_out) return fail(NQA_ERR_INVALID, "null tensor pointer");
  for (int k = 0; k < kNumLevels - 1; ++k)
    if (!feats_out[k]) return fail(NQA_ERR_INVALID, "null feature pointer for level %d", k + 1);
  if ((rc = check_prec(precision)) != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  rc = precision == NQA_PREC_FP32 ? features_impl<float>(ctx, x, B, H, W, precision, feats_out, st)
                                  : features_impl<bf16>(ctx, x, B, H, W, precision, feats_out, st);
  if (rc != NQA_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_dists_from_feats(nqa_ctx* ctx, const float* const feats_x[NQA_NUM_LEVELS],
                         const float* const feats_y[NQA_NUM_LEVELS], int B, int H, int W, float* score_out,
                         float* s1s2_out, void* stream) {
  if (!ctx) return fail(NQA_ERR_INVALID, "null context");
  if (!ctx->have_alpha_beta) return fail(NQA_ERR_STATE, "alpha/beta not set (nqa_set_alpha_beta)");
  if (!feats_x || !feats_y || !score_out) return fail(NQA_ERR_INVALID, "null tensor pointer");
  for (int k = 0; k < kNumLevels; ++k)
    if (!feats_x[k] || !feats_y[k]) return fail(NQA_ERR_INVALID, "null feature pointer for level %d", k);
  int rc = check_shape(B, H, W);
  if (rc != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  PartialBufs pb;
  const size_t need = partial_bytes(B, H, W, NQA_PREC_FP32, &pb, true);
  if ((rc = ensure_arena(ctx, need)) != NQA_OK) return rc;
  ctx->arena.reset();
  partial_alloc(ctx, B, &pb);
  int lh[kNumLevels], lw[kNumLevels];
  level_dims(H, W, lh, lw);
  for (int k = 0; k < kNumLevels; ++k) {
    ProfScope ps(ctx, NQA_K_STATS, st);
    ctx->launches += launch_stats_planar(feats_x[k], feats_y[k], B, level_channels(k), lh[k] * lw[k], pb.chunks[k],
                                         pb.p[k], st);
  }
  rc = finalize(ctx, pb, B, H, W, score_out, s1s2_out, st);
  if (rc != NQA_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

}  // extern "C"

// A-DISTS keeps all six levels of both images alive (the entropy weights are global over the 1475
// channels), as planar fp32 -- the layout its window kernels want and the reference's own
// forward_once format -- so its micro-batch is smaller than the DISTS one.
static size_t adists_feat_bytes(int pairs, int H, int W, size_t elem = sizeof(float)) {
  int h[kNumLevels], w[kNumLevels];
  level_dims(H, W, h, w);
  size_t total = 0;
  for (int k = 1; k < kNumLevels; ++k) total += align256((size_t)2 * pairs * level_channels(k) * h[k] * w[k] * elem);
  return total + 4096;
}

// per-family timing of the A-DISTS stage: adists.cu brackets its two parts through this hook
struct AdistsProfState {
  nqa_ctx* ctx;
  cudaStream_t st;
  ProfScope* open;
};
static AdistsProf adists_prof_hook(AdistsProfState* s) {
  AdistsProf p;
  p.user = s;
  p.begin = [](void* u, int part) {
    auto* ps = static_cast<AdistsProfState*>(u);
    ps->open = new ProfScope(ps->ctx, part == 1 ? NQA_K_ADISTS_WINDOW : NQA_K_ADISTS, ps->st);
  };
  p.end = [](void* u, int launches) {
    auto* ps = static_cast<AdistsProfState*>(u);
    ps->ctx->launches += launches;   // inside the scope, so that the family's launch count is right
    delete ps->open;
    ps->open = nullptr;
  };
  return p;
}

// bf16 trunk: bf16 planes (2 bytes per feature) leave room for the DISTS micro-batch (8 pairs at 1080p: 17 GB of
// workspace); the fp32 / bf16x3 paths carry fp32 planes and take half of it.
static int adists_micro_pairs(const nqa_ctx* ctx, int B, int H, int W, int prec) {
  const int mb = prec == NQA_PREC_BF16 ? pick_micro_pairs(ctx, B, H, W) : pick_micro_pairs(ctx, B, H, W) / 2;
  return mb < 1 ? 1 : mb;
}

// Strip planes (adists.h), bf16 trunk only.  Mode 2 (default): the levels whose producing kernel can write strips get
// them from the convolution epilogue and one kernel pools + sums them (*from_conv); mode 1: strip planes transposed from
// the NHWC level (validation of the readers).  Returns the mask of levels in the strip layout.
static unsigned adists_strip_levels(const nqa_ctx* ctx, int H, int W, int window, bool planes16, unsigned* from_conv) {
  unsigned mask = 0;
  *from_conv = 0;
  if (!planes16 || ctx->adists_strips <= 0 || !ctx->adists_fused_sums) return 0;
  for (int k = 1; k < kNumLevels; ++k) {
    if (!adists_strips_supported(k, H, W, window) || !adists_planes_sums_supported(k, H, W)) continue;
    if (ctx->adists_strips >= 2) {
      if (!trunk_strips_capable(ctx, k)) continue;
      *from_conv |= 1u << k;
    }
    mask |= 1u << k;
  }
  return mask;
}

template <typename T>
static size_t adists_need_bytes(const nqa_ctx* ctx, int mb, int H, int W, int window, int prec, bool with_dists) {
  const bool planes16 = std::is_same<T, bf16>::value && prec == NQA_PREC_BF16;
  const bool fused_pool = with_dists && prec != NQA_PREC_BF16X3;
  unsigned from_conv = 0;
  adists_strip_levels(ctx, H, W, window, planes16, &from_conv);
  return trunk_bytes<T>(2 * mb, H, W, prec) + adists_feat_bytes(mb, H, W, planes16 ? sizeof(bf16) : sizeof(float)) +
         adists_workspace_bytes(mb, H, W, window) +
         (with_dists ? partial_bytes(mb, H, W, prec, nullptr, false, fused_pool, from_conv) : 0);
}

// dists_out != nullptr: the same trunk pass also feeds the DISTS statistics (fused pool + moments kernel), so a
// scorer that wants both metrics for a batch (prep.py:186-190) runs the VGG trunk once instead of twice.
template <typename T>
static int adists_impl(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int window, int prec,
                       float* score_out, float* map_out, cudaStream_t st, float* dists_out = nullptr,
                       const AdistsParts* parts_out = nullptr) {
  const int mb = adists_micro_pairs(ctx, B, H, W, prec);
  const bool with_dists = dists_out != nullptr;
  const bool fused_pool = with_dists && prec != NQA_PREC_BF16X3;
  // bf16 trunk: the planes the statistics read are bf16 too (exact copies of the activations, half the bytes)
  const bool planes16 = std::is_same<T, bf16>::value && prec == NQA_PREC_BF16;
  const size_t pelem = planes16 ? sizeof(bf16) : sizeof(float);
  int rc = ensure_arena(ctx, adists_need_bytes<T>(ctx, mb, H, W, window, prec, with_dists));
  if (rc != NQA_OK) return rc;
  int lh[kNumLevels], lw[kNumLevels];
  level_dims(H, W, lh, lw);
  const size_t frame = (size_t)3 * H * W;
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int pairs = B - b0 < mb ? B - b0 : mb;
    ctx->arena.reset();
    TrunkBufs<T> bufs;
    trunk_alloc<T>(ctx, &bufs, 2 * pairs, H, W, prec);
    char* planar[kNumLevels] = {};
    for (int k = 1; k < kNumLevels; ++k)
      planar[k] = static_cast<char*>(ctx->arena.take((size_t)2 * pairs * level_channels(k) * lh[k] * lw[k] * pelem));
    void* ws = ctx->arena.take(adists_workspace_bytes(pairs, H, W, window));
    unsigned sums_done = 0;   // levels whose channel sums were produced together with the planes (bit 8 + k: from strips)
    // Strip planes (adists.h): written by the convolution epilogues, pooled + summed (+ the DISTS moments when both
    // metrics are wanted) by one kernel; see adists_strip_levels
    unsigned from_conv = 0;
    const unsigned strips_mask = adists_strip_levels(ctx, H, W, window, planes16, &from_conv);
    bf16* strips_dst[kNumLevels] = {};
    for (int k = 1; k < kNumLevels; ++k)
      if ((from_conv >> k) & 1) strips_dst[k] = reinterpret_cast<bf16*>(planar[k]);
    const float* xb = x + b0 * frame;
    const float* yb = y + b0 * frame;
    PartialBufs pb;
    if (with_dists) {
      partial_bytes(pairs, H, W, prec, &pb, false, fused_pool, from_conv);
      partial_alloc(ctx, pairs, &pb);
      ProfScope ps(ctx, NQA_K_STATS, st);
      ctx->launches += launch_stats_planar(xb, yb, pairs, 3, H * W, pb.chunks[0], pb.p[0], st);
    }
    rc = run_trunk<T>(ctx, xb, yb, pairs, 2 * pairs, H, W, prec, bufs, st,
                      [&](int k, const T* f, const T* f_lo, int h, int w, T* pooled_dst) {
                        const int c = level_channels(k);
                        if (strips_dst[k]) {
                          // the level arrived as strip planes: pooled input of the next stage + channel sums in one read
                          ProfScope ps(ctx, NQA_K_L2POOL, st);
                          ctx->launches += adists_pool_strips_and_sums(strips_dst[k], reinterpret_cast<bf16*>(pooled_dst), pairs, k, H, W,
                                                                       window, ws, st, with_dists ? pb.p[k] : nullptr);
                          sums_done |= 1u << (8 + k);
                          return true;
                        }
                        if (planes16 && ctx->adists_fused_sums && adists_planes_sums_supported(k, H, W)) {
                          // transposition + the level's channel sums in one read (adists.cu planes_sums_kernel)
                          ProfScope ps(ctx, NQA_K_LAYOUT, st);
                          ctx->launches += adists_planes_and_sums(reinterpret_cast<const bf16*>(f), reinterpret_cast<bf16*>(planar[k]),
                                                                  pairs, k, H, W, window, ws, st, (strips_mask >> k) & 1);
                          sums_done |= 1u << k;
                        } else if (planes16) {
                          ProfScope ps(ctx, NQA_K_LAYOUT, st);
                          ctx->launches += launch_nhwc_to_nchw_bf16(reinterpret_cast<const bf16*>(f), reinterpret_cast<bf16*>(planar[k]),
                                                                    2 * pairs, c, h, w, st);
                        } else {
                          level_to_nchw<T>(ctx, f, f_lo, reinterpret_cast<float*>(planar[k]), 2 * pairs, c, h, w, st);
                        }
                        if (!with_dists) return false;
                        const size_t half = (size_t)pairs * h * w * c;
                        if (fused_pool && pooled_dst) {
                          const size_t half_out = (size_t)pairs * pooled(h) * pooled(w) * c;
                          ProfScope ps(ctx, NQA_K_POOL_STATS, st);
                          ctx->launches += launch_pool_stats<T>(f, f + half, pooled_dst, pooled_dst + half_out, pairs, h, w, c,
                                                                pb.chunks[k], pb.p[k], st, ctx->pool_stats_tma);
                          return true;
                        }
                        ProfScope ps(ctx, NQA_K_STATS, st);
                        ctx->launches += launch_stats_nhwc<T>(f, f + half, f_lo, f_lo ? f_lo + half : nullptr, pairs, h * w, c,
                                                             pb.chunks[k], pb.p[k], st);
                        return false;
                      }, strips_dst);
    if (rc != NQA_OK) return rc;
    if (with_dists && (rc = finalize(ctx, pb, pairs, H, W, dists_out + b0, nullptr, st)) != NQA_OK) return rc;
    const void* fx[kNumLevels];
    const void* fy[kNumLevels];
    fx[0] = xb;
    fy[0] = yb;
    for (int k = 1; k < kNumLevels; ++k) {
      fx[k] = planar[k];
      fy[k] = planar[k] + (size_t)pairs * level_channels(k) * lh[k] * lw[k] * pelem;
    }
    AdistsParts parts;
    if (parts_out) {   // this micro-batch's slice of the caller's [B, ...] buffers
      if (parts_out->weights) parts.weights = parts_out->weights + (size_t)b0 * kTotalChannels;
      if (parts_out->level_terms) parts.level_terms = parts_out->level_terms + (size_t)b0 * kNumLevels;
      for (int k = 0; k < kNumLevels; ++k)
        if (parts_out->ps[k]) parts.ps[k] = parts_out->ps[k] + (size_t)b0 * adists_ps_elems(H, W, window, k);
    }
    AdistsProfState prof_state{ctx, st, nullptr};
    const AdistsProf prof = adists_prof_hook(&prof_state);
    if (adists_stage(fx, fy, planes16, pairs, H, W, window, ws, score_out + b0, st,
                     map_out ? map_out + (size_t)b0 * H * W : nullptr, parts_out ? &parts : nullptr, &prof,
                     sums_done, &ctx->adists_opt, strips_mask) < 0)   // launches counted by the hook
      return fail(NQA_ERR_CUDA, "A-DISTS statistics stage: no tensor map for the strip planes (%dx%d)", H, W);
  }
  return NQA_OK;
}

static int check_window(int window) {
  if (window < 1 || window > kAdistsMaxWindow || (window & 1) == 0)
    return fail(NQA_ERR_INVALID, "window %d must be odd and in [1,%d]", window, kAdistsMaxWindow);
  if (!adists_window_supported(window))
    return fail(NQA_ERR_INVALID, "this build has window kernels for 7, 11 and 21 (the reference default); got %d", window);
  return NQA_OK;
}

extern "C" {

int nqa_adists_forward(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int window, int precision,
                       float* score_out, void* stream) {
  return nqa_adists_map(ctx, x, y, B, H, W, window, precision, score_out, nullptr, stream);
}

int nqa_adists_map(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int window, int precision,
                   float* score_out, float* map_out, void* stream) {
  int rc = check_ready(ctx, false);
  if (rc != NQA_OK) return rc;
  if (!x || !y || !score_out) return fail(NQA_ERR_INVALID, "null tensor pointer");
  if ((rc = check_prec(precision)) != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  if ((rc = check_window(window)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  auto body = [&](const float* gx, const float* gy, float* const* go, cudaStream_t gs) {
    return precision == NQA_PREC_FP32 ? adists_impl<float>(ctx, gx, gy, B, H, W, window, precision, go[0], map_out, gs)
                                      : adists_impl<bf16>(ctx, gx, gy, B, H, W, window, precision, go[0], map_out, gs);
  };
  float* outs[1] = {score_out};
  rc = map_out ? body(x, y, outs, st) : graphed_call(ctx, 1, x, y, B, H, W, precision, window, outs, 1, st, body);
  if (rc != NQA_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_adists_parts(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int window, int precision,
                     float* score_out, float* weights_out, float* level_terms_out, float* const ps_out[NQA_NUM_LEVELS],
                     void* stream) {
  int rc = check_ready(ctx, false);
  if (rc != NQA_OK) return rc;
  if (!x || !y || !score_out) return fail(NQA_ERR_INVALID, "null tensor pointer");
  if ((rc = check_prec(precision)) != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  if ((rc = check_window(window)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  AdistsParts parts;
  parts.weights = weights_out;
  parts.level_terms = level_terms_out;
  for (int k = 0; k < kNumLevels; ++k) parts.ps[k] = ps_out ? ps_out[k] : nullptr;
  rc = precision == NQA_PREC_FP32
           ? adists_impl<float>(ctx, x, y, B, H, W, window, precision, score_out, nullptr, st, nullptr, &parts)
           : adists_impl<bf16>(ctx, x, y, B, H, W, window, precision, score_out, nullptr, st, nullptr, &parts);
  if (rc != NQA_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_adists_ps_elems(int H, int W, int window, int level, size_t* elems) {
  if (!elems || level < 0 || level >= kNumLevels) return fail(NQA_ERR_INVALID, "bad argument");
  int rc = check_shape(1, H, W);
  if (rc != NQA_OK) return rc;
  if ((rc = check_window(window)) != NQA_OK) return rc;
  *elems = adists_ps_elems(H, W, window, level);
  return NQA_OK;
}

int nqa_dists_adists_forward(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int window, int precision,
                             float* dists_out, float* adists_out, void* stream) {
  int rc = check_ready(ctx, true);
  if (rc != NQA_OK) return rc;
  if (!x || !y || !dists_out || !adists_out) return fail(NQA_ERR_INVALID, "null tensor pointer");
  if ((rc = check_prec(precision)) != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  if ((rc = check_window(window)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  float* outs[2] = {adists_out, dists_out};
  rc = graphed_call(ctx, 2, x, y, B, H, W, precision, window, outs, 2, st,
                    [&](const float* gx, const float* gy, float* const* go, cudaStream_t gs) {
                      return precision == NQA_PREC_FP32
                                 ? adists_impl<float>(ctx, gx, gy, B, H, W, window, precision, go[0], nullptr, gs, go[1])
                                 : adists_impl<bf16>(ctx, gx, gy, B, H, W, window, precision, go[0], nullptr, gs, go[1]);
                    });
  if (rc != NQA_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

}  // extern "C"

// Host-buffer forms of the two A-DISTS entry points (see nqa_dists_forward_host).
static int adists_host_impl(nqa_ctx* ctx, const float* x_host, const float* y_host, int B, int H, int W, int window, int precision,
                            float* dists_host, float* adists_host) {
  const bool with_dists = dists_host != nullptr;
  int rc = check_ready(ctx, with_dists);
  if (rc != NQA_OK) return rc;
  if (!x_host || !y_host || !adists_host) return fail(NQA_ERR_INVALID, "null tensor pointer");
  if ((rc = check_prec(precision)) != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  if ((rc = check_window(window)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  const int mb = adists_micro_pairs(ctx, B, H, W, precision);
  const bool f32 = precision == NQA_PREC_FP32;
  if ((rc = ensure_arena(ctx, f32 ? adists_need_bytes<float>(ctx, mb, H, W, window, precision, with_dists)
                                  : adists_need_bytes<bf16>(ctx, mb, H, W, window, precision, with_dists))) != NQA_OK)
    return rc;
  float* outs[2] = {adists_host, dists_host};
  return host_stream_run(ctx, x_host, y_host, B, H, W, mb, with_dists ? 2 : 1, outs,
                         [&](const float* dx, const float* dy, int pairs, float* const* d_out, cudaStream_t st) {
                           float* dd = with_dists ? d_out[1] : nullptr;
                           return f32 ? adists_impl<float>(ctx, dx, dy, pairs, H, W, window, precision, d_out[0], nullptr, st, dd)
                                      : adists_impl<bf16>(ctx, dx, dy, pairs, H, W, window, precision, d_out[0], nullptr, st, dd);
                         });
}

extern "C" {

int nqa_adists_forward_host(nqa_ctx* ctx, const float* x_host, const float* y_host, int B, int H, int W, int window,
                            int precision, float* score_host) {
  return adists_host_impl(ctx, x_host, y_host, B, H, W, window, precision, nullptr, score_host);
}

int nqa_dists_adists_forward_host(nqa_ctx* ctx, const float* x_host, const float* y_host, int B, int H, int W, int window,
                                  int precision, float* dists_host, float* adists_host) {
  if (!dists_host) return fail(NQA_ERR_INVALID, "null tensor pointer");
  return adists_host_impl(ctx, x_host, y_host, B, H, W, window, precision, dists_host, adists_host);
}

int nqa_adists_from_feats(nqa_ctx* ctx, const float* const feats_x[NQA_NUM_LEVELS],
                          const float* const feats_y[NQA_NUM_LEVELS], int B, int H, int W, int window, float* score_out,
                          void* stream) {
  if (!ctx) return fail(NQA_ERR_INVALID, "null context");
  if (!feats_x || !feats_y || !score_out) return fail(NQA_ERR_INVALID, "null tensor pointer");
  for (int k = 0; k < kNumLevels; ++k)
    if (!feats_x[k] || !feats_y[k]) return fail(NQA_ERR_INVALID, "null feature pointer for level %d", k);
  int rc = check_shape(B, H, W);
  if (rc != NQA_OK) return rc;
  if ((rc = check_window(window)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  if ((rc = ensure_arena(ctx, adists_workspace_bytes(B, H, W, window) + 4096)) != NQA_OK) return rc;
  ctx->arena.reset();
  void* ws = ctx->arena.take(adists_workspace_bytes(B, H, W, window));
  {
    AdistsProfState prof_state{ctx, st, nullptr};
    const AdistsProf prof = adists_prof_hook(&prof_state);
    const void* vx[kNumLevels];
    const void* vy[kNumLevels];
    for (int k = 0; k < kNumLevels; ++k) { vx[k] = feats_x[k]; vy[k] = feats_y[k]; }
    adists_stage(vx, vy, false, B, H, W, window, ws, score_out, st, nullptr, nullptr, &prof, 0, &ctx->adists_opt);   // launches counted by the hook
  }
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_dists_backward(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int precision,
                       const float* grad_score, int through_trunk, float* grad_x, float* grad_y, void* stream) {
  int rc = check_ready(ctx, true);
  if (rc != NQA_OK) return rc;
  if (!x || !y || !grad_score || !grad_x || !grad_y) return fail(NQA_ERR_INVALID, "null tensor pointer");
  if ((rc = check_prec(precision)) != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  int mb = pick_micro_pairs(ctx, B, H, W) / (precision == NQA_PREC_BF16X3 ? 4 : 2);
  if (mb < 1) mb = 1;
  const size_t need = precision == NQA_PREC_FP32 ? dists_bwd_bytes<float>(ctx, mb, H, W, precision)
                                                 : dists_bwd_bytes<bf16>(ctx, mb, H, W, precision);
  if ((rc = ensure_arena(ctx, need)) != NQA_OK) return rc;
  const size_t frame = (size_t)3 * H * W;
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int pairs = B - b0 < mb ? B - b0 : mb;
    rc = precision == NQA_PREC_FP32
             ? dists_backward_micro_batch<float>(ctx, x + b0 * frame, y + b0 * frame, pairs, H, W, precision, grad_score + b0,
                                                 through_trunk != 0, grad_x + b0 * frame, grad_y + b0 * frame, st)
             : dists_backward_micro_batch<bf16>(ctx, x + b0 * frame, y + b0 * frame, pairs, H, W, precision, grad_score + b0,
                                                through_trunk != 0, grad_x + b0 * frame, grad_y + b0 * frame, st);
    if (rc != NQA_OK) return rc;
  }
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_dists_forward_saved(nqa_ctx* ctx, const float* x, const float* y, int B, int H, int W, int precision,
                            float* score_out, float* s1s2_out, unsigned long long* tape_out, void* stream) {
  int rc = check_ready(ctx, true);
  if (rc != NQA_OK) return rc;
  if (!x || !y || !score_out || !tape_out) return fail(NQA_ERR_INVALID, "null tensor pointer");
  if ((rc = check_prec(precision)) != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  *tape_out = 0;
  int mb = pick_micro_pairs(ctx, B, H, W) / (precision == NQA_PREC_BF16X3 ? 4 : 2);
  if (mb < 1) mb = 1;
  const bool f32 = precision == NQA_PREC_FP32;
  size_t tape_bytes = 0;
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int pairs = B - b0 < mb ? B - b0 : mb;
    tape_bytes += align256(f32 ? saved_bytes<float>(pairs, H, W, precision) : saved_bytes<bf16>(pairs, H, W, precision));
  }
  const int buf = acquire_tape_buf(ctx, tape_bytes);
  if (buf < 0)   // does not fit: plain forward, the caller differentiates by recomputation (nqa_dists_backward)
    return dists_forward_device(ctx, x, y, B, H, W, precision, score_out, s1s2_out, st);
  const size_t need = partial_bytes(mb, H, W, precision, nullptr, false, precision != NQA_PREC_BF16X3) + 8192;
  if ((rc = ensure_arena(ctx, need)) != NQA_OK) {
    ctx->tape_bufs[buf].in_use = false;
    return rc;
  }
  Tape tape;
  tape.id = ctx->next_tape++;
  tape.buf = buf;
  tape.B = B; tape.H = H; tape.W = W; tape.prec = precision; tape.mb = mb;
  Arena store;
  store.base = ctx->tape_bufs[buf].base;
  store.cap = ctx->tape_bufs[buf].cap;
  const size_t frame = (size_t)3 * H * W;
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int pairs = B - b0 < mb ? B - b0 : mb;
    ctx->arena.reset();
    SavedActs sv;
    float* s12 = s1s2_out ? s1s2_out + (size_t)b0 * 2 * kTotalChannels : nullptr;
    rc = f32 ? trunk_forward_keep<float>(ctx, store, x + b0 * frame, y + b0 * frame, pairs, H, W, precision, true, &sv,
                                         score_out + b0, s12, st)
             : trunk_forward_keep<bf16>(ctx, store, x + b0 * frame, y + b0 * frame, pairs, H, W, precision, true, &sv,
                                        score_out + b0, s12, st);
    if (rc != NQA_OK) {
      ctx->tape_bufs[buf].in_use = false;
      return rc;
    }
    tape.chunks.push_back(sv);
  }
  if (store.used > store.cap) {   // cannot happen (saved_bytes is the sum of the same takes); refuse to hand out the tape
    ctx->tape_bufs[buf].in_use = false;
    return fail(NQA_ERR_STATE, "activation tape overflow (%zu > %zu bytes)", store.used, store.cap);
  }
  CUDA_TRY(cudaGetLastError());
  ctx->tapes.push_back(tape);
  *tape_out = tape.id;
  return NQA_OK;
}

int nqa_dists_backward_saved(nqa_ctx* ctx, unsigned long long tape_id, const float* x, const float* y, const float* grad_score,
                             float* grad_x, float* grad_y, void* stream) {
  int rc = check_ready(ctx, true);
  if (rc != NQA_OK) return rc;
  if (!x || !y || !grad_score || !grad_x || !grad_y) return fail(NQA_ERR_INVALID, "null tensor pointer");
  Tape* tp = find_tape(ctx, tape_id);
  if (!tp) return fail(NQA_ERR_STATE, "unknown or already consumed activation tape %llu", tape_id);
  const Tape tape = *tp;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  const bool f32 = tape.prec == NQA_PREC_FP32;
  const int B = tape.B, H = tape.H, W = tape.W, mb = tape.mb;
  const size_t need = f32 ? bwd_work_bytes<float>(mb, H, W, tape.prec) : bwd_work_bytes<bf16>(mb, H, W, tape.prec);
  if ((rc = ensure_arena(ctx, need)) != NQA_OK) return rc;
  const size_t frame = (size_t)3 * H * W;
  int chunk = 0;
  for (int b0 = 0; b0 < B; b0 += mb, ++chunk) {
    const int pairs = B - b0 < mb ? B - b0 : mb;
    ctx->arena.reset();
    const SavedActs& sv = tape.chunks[chunk];
    rc = f32 ? trunk_backward_walk<float>(ctx, x + b0 * frame, y + b0 * frame, pairs, H, W, sv, grad_score + b0, true,
                                          grad_x + b0 * frame, grad_y + b0 * frame, st)
             : trunk_backward_walk<bf16>(ctx, x + b0 * frame, y + b0 * frame, pairs, H, W, sv, grad_score + b0, true,
                                         grad_x + b0 * frame, grad_y + b0 * frame, st);
    if (rc != NQA_OK) break;
  }
  release_tape(ctx, tape_id);   // the buffer is reused by later calls on the same stream order
  if (rc != NQA_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_tape_release(nqa_ctx* ctx, unsigned long long tape_id) {
  if (!ctx) return fail(NQA_ERR_INVALID, "null context");
  release_tape(ctx, tape_id);
  return NQA_OK;
}

int nqa_dists_from_feats_backward(nqa_ctx* ctx, const float* const feats_x[NQA_NUM_LEVELS],
                                  const float* const feats_y[NQA_NUM_LEVELS], int B, int H, int W, const float* grad_score,
                                  float* const grad_fx[NQA_NUM_LEVELS], float* const grad_fy[NQA_NUM_LEVELS], void* stream) {
  if (!ctx) return fail(NQA_ERR_INVALID, "null context");
  if (!ctx->have_alpha_beta) return fail(NQA_ERR_STATE, "alpha/beta not set (nqa_set_alpha_beta)");
  if (!feats_x || !feats_y || !grad_score || !grad_fx || !grad_fy) return fail(NQA_ERR_INVALID, "null tensor pointer");
  for (int k = 0; k < kNumLevels; ++k)
    if (!feats_x[k] || !feats_y[k] || !grad_fx[k] || !grad_fy[k]) return fail(NQA_ERR_INVALID, "null pointer for level %d", k);
  int rc = check_shape(B, H, W);
  if (rc != NQA_OK) return rc;
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  PartialBufs pb;
  const size_t need = partial_bytes(B, H, W, NQA_PREC_FP32, &pb, true) + align256((size_t)B * kTotalChannels * 5 * sizeof(float)) +
                      align256((size_t)B * 5 * kCoefStride * sizeof(float)) + align256(B * sizeof(float)) + 4096;
  if ((rc = ensure_arena(ctx, need)) != NQA_OK) return rc;
  ctx->arena.reset();
  partial_alloc(ctx, B, &pb);
  float* moments = static_cast<float*>(ctx->arena.take((size_t)B * kTotalChannels * 5 * sizeof(float)));
  float* coef = static_cast<float*>(ctx->arena.take((size_t)B * 5 * kCoefStride * sizeof(float)));
  float* score_scratch = static_cast<float*>(ctx->arena.take(B * sizeof(float)));
  int lh[kNumLevels], lw[kNumLevels];
  level_dims(H, W, lh, lw);
  GradLevels gl;
  for (int k = 0; k < kNumLevels; ++k) {
    gl.inv_n[k] = 1.0 / ((double)lh[k] * lw[k]);
    ProfScope ps(ctx, NQA_K_STATS, st);
    ctx->launches += launch_stats_planar(feats_x[k], feats_y[k], B, level_channels(k), lh[k] * lw[k], pb.chunks[k], pb.p[k], st);
  }
  rc = finalize(ctx, pb, B, H, W, score_scratch, nullptr, st, moments);
  if (rc != NQA_OK) return rc;
  {
    ProfScope ps(ctx, NQA_K_BACKWARD, st);
    ctx->launches += launch_stat_coef(moments, ctx->alpha, ctx->beta, ctx->ab_normalize, grad_score, gl, kNumLevels, B, coef, st);
    for (int k = 0; k < kNumLevels; ++k)
      ctx->launches += launch_stat_grad_planar(feats_x[k], feats_y[k], coef, k, B, level_channels(k), lh[k] * lw[k],
                                               grad_fx[k], grad_fy[k], false, st);
  }
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_l2pool(nqa_ctx* ctx, const float* x, int B, int C, int H, int W, float* out, void* stream) {
  if (!ctx || !x || !out) return fail(NQA_ERR_INVALID, "null argument");
  if (B < 1 || C < 1 || H < 1 || W < 1) return fail(NQA_ERR_SHAPE, "empty tensor %dx%dx%dx%d", B, C, H, W);
  DeviceGuard guard(ctx->device);
  ctx->launches += launch_l2pool_nchw(x, out, B, C, H, W, static_cast<cudaStream_t>(stream));
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_resize_bilinear(nqa_ctx* ctx, const float* x, int B, int C, int H, int W, int OH, int OW, int antialias, float* out,
                        void* stream) {
  if (!ctx || !x || !out) return fail(NQA_ERR_INVALID, "null argument");
  if (B < 1 || C < 1 || H < 1 || W < 1 || OH < 1 || OW < 1)
    return fail(NQA_ERR_SHAPE, "empty tensor %dx%dx%dx%d -> %dx%d", B, C, H, W, OH, OW);
  DeviceGuard guard(ctx->device);
  ctx->launches += launch_resize_bilinear(x, out, B * C, H, W, OH, OW, antialias != 0, static_cast<cudaStream_t>(stream));
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_conv3x3_relu(nqa_ctx* ctx, const float* x, const float* w, const float* bias, int B, int Cin, int Cout, int H,
                     int W, int precision, float* out, void* stream) {
  if (!ctx || !x || !w || !bias || !out) return fail(NQA_ERR_INVALID, "null argument");
  int rc = check_prec(precision);
  if (rc != NQA_OK) return rc;
  if ((rc = check_shape(B, H, W)) != NQA_OK) return rc;
  if (Cin % 64 != 0 || Cout % 64 != 0 || Cin < 64 || Cout < 64 || Cout > 512 || Cin > 512)
    return fail(NQA_ERR_SHAPE, "Cin/Cout must be multiples of 64 in [64,512], got %d/%d", Cin, Cout);
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StreamOrder so(ctx, st);
  if (so.rc != NQA_OK) return so.rc;
  const size_t nin = (size_t)B * H * W * Cin, nout = (size_t)B * H * W * Cout, nw = (size_t)Cout * Cin * 9;
  const size_t need = 2 * align256(nin * 4) + 2 * align256(nout * 4) + align256(nw * 4) + 2 * align256(nw * 2) + 8192;
  if ((rc = ensure_arena(ctx, need)) != NQA_OK) return rc;
  ctx->arena.reset();
  float* w_f32 = static_cast<float*>(ctx->arena.take(nw * 4));
  bf16* w_hi = static_cast<bf16*>(ctx->arena.take(nw * 2));
  bf16* w_lo = static_cast<bf16*>(ctx->arena.take(nw * 2));
  pack_w_kernel<<<1024, 256, 0, st>>>(w, Cin, Cout, w_f32, w_hi, w_lo);
  ctx->launches += 1;
  if (precision == NQA_PREC_FP32) {
    float* in_t = static_cast<float*>(ctx->arena.take(nin * 4));
    float* out_t = static_cast<float*>(ctx->arena.take(nout * 4));
    ctx->launches += launch_nchw_to_nhwc<float>(x, in_t, nullptr, B, Cin, H, W, st);
    ctx->launches += launch_conv3x3_f32(in_t, out_t, w_f32, bias, B, H, W, Cin, Cout, st);
    ctx->launches += launch_nhwc_to_nchw<float>(out_t, out, B, Cout, H, W, st);
  } else {
    const bool split = precision == NQA_PREC_BF16X3;
    bf16* in_hi = static_cast<bf16*>(ctx->arena.take(nin * 2));
    bf16* in_lo = static_cast<bf16*>(ctx->arena.take(nin * 2));
    bf16* out_hi = static_cast<bf16*>(ctx->arena.take(nout * 2));
    bf16* out_lo = static_cast<bf16*>(ctx->arena.take(nout * 2));
    ctx->launches += launch_nchw_to_nhwc<bf16>(x, in_hi, split ? in_lo : nullptr, B, Cin, H, W, st);
    ConvTcPlan plan;
    const char* err = "";
    if (conv_tc_make_plan(&plan, ctx->conv_cfg, in_hi, in_lo, w_hi, w_lo, bias, out_hi, out_lo, B, H, W, Cin, Cout, split, &err,
                          false, /*split_on_pairs=*/split))
      return fail(NQA_ERR_CUDA, "%s", err);
    const int nl = launch_conv3x3_tc(plan, st);
    if (nl <= 0) return fail(NQA_ERR_CUDA, "convolution could not be launched (%dx%dx%d, %d -> %d channels)", B, H, W, Cin, Cout);
    ctx->launches += nl;
    if (split)
      ctx->launches += launch_nhwc_split_to_nchw(out_hi, out_lo, out, B, Cout, H, W, st);
    else
      ctx->launches += launch_nhwc_to_nchw<bf16>(out_hi, out, B, Cout, H, W, st);
  }
  CUDA_TRY(cudaGetLastError());
  return NQA_OK;
}

int nqa_workspace_bytes(nqa_ctx* ctx, size_t* bytes) {
  if (!ctx || !bytes) return fail(NQA_ERR_INVALID, "null argument");
  *bytes = ctx->arena.cap + 2 * ctx->stage_cap * sizeof(float);
  return NQA_OK;
}

int nqa_launch_count(nqa_ctx* ctx, uint64_t* launches) {
  if (!ctx || !launches) return fail(NQA_ERR_INVALID, "null argument");
  *launches = ctx->launches;
  return NQA_OK;
}

int nqa_graph_replays(nqa_ctx* ctx, uint64_t* replays) {
  if (!ctx || !replays) return fail(NQA_ERR_INVALID, "null argument");
  *replays = ctx->graph_replays;
  return NQA_OK;
}

int nqa_profile_enable(nqa_ctx* ctx, int on) {
  if (!ctx) return fail(NQA_ERR_INVALID, "null context");
  ctx->profiling = on != 0;
  return NQA_OK;
}

int nqa_profile_read(nqa_ctx* ctx, double ms[NQA_NUM_KERNEL_FAMILIES], uint64_t launches[NQA_NUM_KERNEL_FAMILIES]) {
  if (!ctx || !ms || !launches) return fail(NQA_ERR_INVALID, "null argument");
  DeviceGuard guard(ctx->device);
  CUDA_TRY(cudaDeviceSynchronize());
  for (int i = 0; i < NQA_NUM_KERNEL_FAMILIES; ++i) { ms[i] = 0.0; launches[i] = 0; }
  for (auto& r : ctx->prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.family >= 0 && r.family < NQA_NUM_KERNEL_FAMILIES) {
      ms[r.family] += (double)t;
      launches[r.family] += r.launches;
    }
    ctx->prof_pool.push_back(r.a);
    ctx->prof_pool.push_back(r.b);
  }
  ctx->prof_recs.clear();
  cudaGetLastError();
  return NQA_OK;
}

int nqa_measure_rate(nqa_ctx* ctx, int which, double* fma_per_s) {
  if (!ctx || !fma_per_s) return fail(NQA_ERR_INVALID, "null argument");
  if (which != NQA_RATE_FP32_FMA && which != NQA_RATE_MMA_SYNC_BF16) return fail(NQA_ERR_INVALID, "unknown rate %d", which);
  DeviceGuard guard(ctx->device);
  if (microbench_rate(which, ctx->sm_count, ctx->alpha /* never written */, nullptr, fma_per_s) != 0) {
    cudaGetLastError();
    return fail(NQA_ERR_CUDA, "micro-benchmark failed");
  }
  return NQA_OK;
}

int nqa_micro_batch(nqa_ctx* ctx, int B, int H, int W, int precision, int* pairs) {
  if (!ctx || !pairs) return fail(NQA_ERR_INVALID, "null argument");
  int rc = check_shape(B, H, W);
  if (rc != NQA_OK) return rc;
  (void)precision;
  *pairs = pick_micro_pairs(ctx, B, H, W);
  return NQA_OK;
}

}  // extern "C"
