// Thin PyTorch C++ extension over the C-ABI library (include/nerfqa_b200.h): the scoring entry points as torch custom
// operators, torch.ops.nerfqa_b200.*.  It owns no kernels and no arithmetic -- every op checks its tensors, allocates the
// outputs with torch's allocator, takes torch's current CUDA stream and calls one nqa_* function with raw pointers.  The
// nn.Module classes (nerf_qa_b200/dists.py, adists.py) reach the library through these ops when the extension is built
// (nerf_qa_b200/_lib.py), so that the dispatcher, torch.compile / export (fake kernels are registered in torch_ops.py)
// and C++ callers of the dispatcher see the same operators a Python caller does.
//
// Reference call sites these ops stand behind: DISTS.forward (nerf_qa/DISTS_pytorch/DISTS_pt.py:105-148), forward_once
// (:91-103), ADISTS.forward (nerf_qa/ADISTS/ADISTS.py:137-197), L2pooling.forward (DISTS_pt.py:22-25) and the pair of calls
// of prep.py:186-190.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>
#include <vector>

#include "nerfqa_b200.h"

namespace {

nqa_ctx* ctx_of(int64_t handle) {
  TORCH_CHECK(handle != 0, "nerfqa_b200: null context handle");
  return reinterpret_cast<nqa_ctx*>(static_cast<intptr_t>(handle));
}

void check(int rc) { TORCH_CHECK(rc == NQA_OK, "nerfqa_b200 error ", rc, ": ", nqa_last_error()); }

// fp32, contiguous, on a CUDA device: there is no CPU path behind these operators
at::Tensor frames(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor: nerf_qa_b200 is B200-only and has no CPU fallback");
  TORCH_CHECK(t.dim() == 4 && t.size(1) == 3, "expected [B,3,H,W] frames, got ", t.sizes());
  return t.detach().to(at::kFloat).contiguous();
}

void* stream_of(const at::Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

int64_t ctx_create(int64_t device) {
  nqa_ctx* ctx = nullptr;
  check(nqa_ctx_create(static_cast<int>(device), &ctx));
  return static_cast<int64_t>(reinterpret_cast<intptr_t>(ctx));
}

void ctx_destroy(int64_t handle) { check(nqa_ctx_destroy(ctx_of(handle))); }

// (score [B], s1s2 [B,2,1475]): DISTS.forward + the per-channel S1 / S2 the alpha / beta gradients need
std::tuple<at::Tensor, at::Tensor> dists_forward(int64_t handle, const at::Tensor& x_, const at::Tensor& y_, int64_t precision,
                                                 bool want_s1s2) {
  TORCH_CHECK(x_.sizes() == y_.sizes(), "shape mismatch: x ", x_.sizes(), " vs y ", y_.sizes());
  const at::Tensor x = frames(x_, "x"), y = frames(y_, "y");
  const c10::cuda::CUDAGuard guard(x.device());
  const int64_t b = x.size(0);
  at::Tensor score = at::empty({b}, x.options());
  at::Tensor s1s2 = at::empty({want_s1s2 ? b : 0, 2, NQA_NUM_CHANNELS}, x.options());
  check(nqa_dists_forward(ctx_of(handle), x.data_ptr<float>(), y.data_ptr<float>(), (int)b, (int)x.size(2), (int)x.size(3),
                          (int)precision, score.data_ptr<float>(), want_s1s2 ? s1s2.data_ptr<float>() : nullptr, stream_of(x)));
  return {score, s1s2};
}

at::Tensor adists_forward(int64_t handle, const at::Tensor& x_, const at::Tensor& y_, int64_t window, int64_t precision) {
  TORCH_CHECK(x_.sizes() == y_.sizes(), "shape mismatch: x ", x_.sizes(), " vs y ", y_.sizes());
  const at::Tensor x = frames(x_, "x"), y = frames(y_, "y");
  const c10::cuda::CUDAGuard guard(x.device());
  at::Tensor score = at::empty({x.size(0)}, x.options());
  check(nqa_adists_forward(ctx_of(handle), x.data_ptr<float>(), y.data_ptr<float>(), (int)x.size(0), (int)x.size(2),
                           (int)x.size(3), (int)window, (int)precision, score.data_ptr<float>(), stream_of(x)));
  return score;
}

// (DISTS [B], A-DISTS [B]) from one trunk evaluation
std::tuple<at::Tensor, at::Tensor> dists_adists_forward(int64_t handle, const at::Tensor& x_, const at::Tensor& y_,
                                                        int64_t window, int64_t precision) {
  TORCH_CHECK(x_.sizes() == y_.sizes(), "shape mismatch: x ", x_.sizes(), " vs y ", y_.sizes());
  const at::Tensor x = frames(x_, "x"), y = frames(y_, "y");
  const c10::cuda::CUDAGuard guard(x.device());
  at::Tensor d = at::empty({x.size(0)}, x.options()), a = at::empty({x.size(0)}, x.options());
  check(nqa_dists_adists_forward(ctx_of(handle), x.data_ptr<float>(), y.data_ptr<float>(), (int)x.size(0), (int)x.size(2),
                                 (int)x.size(3), (int)window, (int)precision, d.data_ptr<float>(), a.data_ptr<float>(),
                                 stream_of(x)));
  return {d, a};
}

// levels 1..5 of forward_once as NCHW fp32 tensors (level 0 is x itself)
std::vector<at::Tensor> dists_features(int64_t handle, const at::Tensor& x_, int64_t precision) {
  static const int64_t chns[NQA_NUM_LEVELS - 1] = {64, 128, 256, 512, 512};
  const at::Tensor x = frames(x_, "x");
  const c10::cuda::CUDAGuard guard(x.device());
  std::vector<at::Tensor> outs;
  float* ptrs[NQA_NUM_LEVELS - 1];
  int64_t h = x.size(2), w = x.size(3);
  for (int k = 0; k < NQA_NUM_LEVELS - 1; ++k) {
    if (k > 0) { h = (h - 1) / 2 + 1; w = (w - 1) / 2 + 1; }
    outs.push_back(at::empty({x.size(0), chns[k], h, w}, x.options()));
    ptrs[k] = outs.back().data_ptr<float>();
  }
  check(nqa_dists_features(ctx_of(handle), x.data_ptr<float>(), (int)x.size(0), (int)x.size(2), (int)x.size(3), (int)precision,
                           ptrs, stream_of(x)));
  return outs;
}

at::Tensor l2pool(int64_t handle, const at::Tensor& x_) {
  TORCH_CHECK(x_.is_cuda() && x_.dim() == 4, "l2pool expects a CUDA [B,C,H,W] tensor");
  const at::Tensor x = x_.detach().to(at::kFloat).contiguous();
  const c10::cuda::CUDAGuard guard(x.device());
  at::Tensor out = at::empty({x.size(0), x.size(1), (x.size(2) - 1) / 2 + 1, (x.size(3) - 1) / 2 + 1}, x.options());
  check(nqa_l2pool(ctx_of(handle), x.data_ptr<float>(), (int)x.size(0), (int)x.size(1), (int)x.size(2), (int)x.size(3),
                   out.data_ptr<float>(), stream_of(x)));
  return out;
}

}  // namespace

TORCH_LIBRARY(nerfqa_b200, m) {
  m.def("ctx_create(int device) -> int", &ctx_create);
  m.def("ctx_destroy(int ctx) -> ()", &ctx_destroy);
  m.def("dists_forward(int ctx, Tensor x, Tensor y, int precision, bool want_s1s2) -> (Tensor, Tensor)");
  m.def("adists_forward(int ctx, Tensor x, Tensor y, int window, int precision) -> Tensor");
  m.def("dists_adists_forward(int ctx, Tensor x, Tensor y, int window, int precision) -> (Tensor, Tensor)");
  m.def("dists_features(int ctx, Tensor x, int precision) -> Tensor[]");
  m.def("l2pool(int ctx, Tensor x) -> Tensor");
}

TORCH_LIBRARY_IMPL(nerfqa_b200, CUDA, m) {
  m.impl("dists_forward", &dists_forward);
  m.impl("adists_forward", &adists_forward);
  m.impl("dists_adists_forward", &dists_adists_forward);
  m.impl("dists_features", &dists_features);
  m.impl("l2pool", &l2pool);
}
