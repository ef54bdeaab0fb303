"""Binding of the C-ABI library ``libnerfqa_b200.so`` (include/nerfqa_b200.h).

PyTorch is plumbing here (device memory, streams); every scoring call goes through the C-ABI with raw
pointers.  The no-grad scoring calls (``dists_forward``, ``adists_forward``, ``dists_adists_forward``, ``features``,
``l2pool``) go through the thin torch extension ``torch.ops.nerfqa_b200.*`` (torch_ops.py, csrc_torch/) when it is built,
everything else -- and everything when ``NQA_BINDING=ctypes`` -- through ctypes; both reach the same ``nqa_*`` entry points.
There is no CPU fallback: if the library is missing or no CUDA device exists, calls raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnerfqa_b200.so")

PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3}
NUM_CONVS, NUM_LEVELS, NUM_CHANNELS = 13, 6, 1475
CHNS = (3, 64, 128, 256, 512, 512)

EXPORTS = (
    "nqa_last_error", "nqa_version", "nqa_ctx_create", "nqa_ctx_destroy", "nqa_set_vgg_weights",
    "nqa_set_alpha_beta", "nqa_dists_forward", "nqa_dists_forward_host", "nqa_dists_features",
    "nqa_dists_from_feats", "nqa_adists_forward", "nqa_adists_from_feats", "nqa_l2pool",
    "nqa_workspace_bytes", "nqa_launch_count", "nqa_micro_batch", "nqa_conv3x3_relu",
    "nqa_profile_enable", "nqa_profile_read", "nqa_dists_backward", "nqa_dists_from_feats_backward",
    "nqa_set_channel_weights", "nqa_adists_map", "nqa_dists_adists_forward", "nqa_dists_forward_saved",
    "nqa_dists_backward_saved", "nqa_tape_release", "nqa_set_alpha_beta_async", "nqa_set_channel_weights_async",
    "nqa_adists_forward_host", "nqa_dists_adists_forward_host", "nqa_adists_parts", "nqa_adists_ps_elems",
    "nqa_measure_rate", "nqa_resize_bilinear", "nqa_graph_replays",
)
KERNEL_FAMILIES = ("conv_tc", "conv1_1", "conv_f32", "l2pool", "stats", "finalize", "layout", "adists", "pool_stats",
                   "stage1", "backward", "adists_window")


class NerfQAError(RuntimeError):
    pass


_lib: Optional[ctypes.CDLL] = None


def load_library() -> ctypes.CDLL:
    """Load (once) and type the shared library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NerfQAError(
            f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "nerf_qa_b200 has no CPU / eager-PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, fp = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p
    lib.nqa_last_error.restype = ctypes.c_char_p
    lib.nqa_last_error.argtypes = []
    lib.nqa_version.restype = ci
    lib.nqa_ctx_create.argtypes = [ci, ctypes.POINTER(vp)]
    lib.nqa_ctx_destroy.argtypes = [vp]
    lib.nqa_set_vgg_weights.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    lib.nqa_set_alpha_beta.argtypes = [vp, fp, fp]
    lib.nqa_set_channel_weights.argtypes = [vp, fp, fp]
    lib.nqa_set_alpha_beta_async.argtypes = [vp, fp, fp, vp]
    lib.nqa_set_channel_weights_async.argtypes = [vp, fp, fp, vp]
    lib.nqa_adists_forward_host.argtypes = [vp, fp, fp, ci, ci, ci, ci, ci, fp]
    lib.nqa_dists_adists_forward_host.argtypes = [vp, fp, fp, ci, ci, ci, ci, ci, fp, fp]
    lib.nqa_dists_forward.argtypes = [vp, fp, fp, ci, ci, ci, ci, fp, fp, vp]
    lib.nqa_dists_forward_host.argtypes = [vp, fp, fp, ci, ci, ci, ci, fp]
    lib.nqa_dists_features.argtypes = [vp, fp, ci, ci, ci, ci, ctypes.POINTER(vp), vp]
    lib.nqa_dists_from_feats.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, ci, fp, fp, vp]
    lib.nqa_adists_forward.argtypes = [vp, fp, fp, ci, ci, ci, ci, ci, fp, vp]
    lib.nqa_dists_adists_forward.argtypes = [vp, fp, fp, ci, ci, ci, ci, ci, fp, fp, vp]
    lib.nqa_adists_parts.argtypes = [vp, fp, fp, ci, ci, ci, ci, ci, fp, fp, fp, ctypes.POINTER(vp), vp]
    lib.nqa_adists_ps_elems.argtypes = [ci, ci, ci, ci, ctypes.POINTER(ctypes.c_size_t)]
    lib.nqa_adists_map.argtypes = [vp, fp, fp, ci, ci, ci, ci, ci, fp, fp, vp]
    lib.nqa_adists_from_feats.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, ci, ci, fp, vp]
    lib.nqa_l2pool.argtypes = [vp, fp, ci, ci, ci, ci, fp, vp]
    lib.nqa_dists_backward.argtypes = [vp, fp, fp, ci, ci, ci, ci, fp, ci, fp, fp, vp]
    lib.nqa_dists_forward_saved.argtypes = [vp, fp, fp, ci, ci, ci, ci, fp, fp, ctypes.POINTER(ctypes.c_ulonglong), vp]
    lib.nqa_dists_backward_saved.argtypes = [vp, ctypes.c_ulonglong, fp, fp, fp, fp, fp, vp]
    lib.nqa_tape_release.argtypes = [vp, ctypes.c_ulonglong]
    lib.nqa_dists_from_feats_backward.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, ci, fp,
                                                  ctypes.POINTER(vp), ctypes.POINTER(vp), vp]
    lib.nqa_conv3x3_relu.argtypes = [vp, fp, fp, fp, ci, ci, ci, ci, ci, ci, fp, vp]
    lib.nqa_workspace_bytes.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t)]
    lib.nqa_launch_count.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
    lib.nqa_graph_replays.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
    lib.nqa_profile_enable.argtypes = [vp, ci]
    lib.nqa_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    lib.nqa_resize_bilinear.argtypes = [vp, fp, ci, ci, ci, ci, ci, ci, ci, fp, vp]
    lib.nqa_measure_rate.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_double)]
    lib.nqa_micro_batch.argtypes = [vp, ci, ci, ci, ci, ctypes.POINTER(ci)]
    for name in EXPORTS:
        if name not in ("nqa_last_error",):
            getattr(lib, name).restype = ci
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        msg = load_library().nqa_last_error().decode("utf-8", "replace")
        raise NerfQAError(f"nerfqa_b200 error {rc}: {msg}")


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def _stream(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _as_input(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise NerfQAError(f"{name} must be a CUDA tensor: nerf_qa_b200 is B200-only and has no CPU fallback")
    return t.detach().to(torch.float32).contiguous()


def level_shapes(b: int, h: int, w: int):
    shapes = [(b, 3, h, w), (b, 64, h, w)]
    for c in CHNS[2:]:
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        shapes.append((b, c, h, w))
    return shapes


class ActivationTape:
    """Handle of the activations ``Context.dists_forward_saved`` kept in the library; released when it is consumed by
    ``dists_backward_saved`` or garbage-collected (a forward that is never differentiated)."""

    def __init__(self, ctx: "Context", tape_id: int):
        self.ctx, self.id = ctx, tape_id

    def release(self):
        if self.id and self.ctx.handle is not None:
            self.ctx.lib.nqa_tape_release(self.ctx.handle, self.id)
        self.id = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Context:
    """One library context per (module, GPU)."""

    def __init__(self, device: torch.device):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise NerfQAError("no CUDA device: nerf_qa_b200 has no CPU fallback")
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        h = ctypes.c_void_p()
        _check(self.lib.nqa_ctx_create(self.device.index, ctypes.byref(h)))
        self.handle = h
        self.weights_key = None
        self.ab_key = None
        from . import torch_ops
        self.ops = torch_ops.load_if_enabled()   # torch.ops.nerfqa_b200 or None (ctypes only)

    def _op(self, fn, *args):
        """One torch.ops.nerfqa_b200 call on this context; library errors surface as NerfQAError like the ctypes path."""
        try:
            return fn(int(self.handle.value), *args)
        except RuntimeError as e:
            raise NerfQAError(str(e)) from None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.nqa_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- parameters ------------------------------------------------------------------------------
    def set_vgg_weights(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]) -> None:
        assert len(weights) == NUM_CONVS and len(biases) == NUM_CONVS
        ws = [w.detach().to(torch.float32).contiguous() for w in weights]
        bs = [b.detach().to(torch.float32).contiguous() for b in biases]
        wa = (ctypes.c_void_p * NUM_CONVS)(*[w.data_ptr() for w in ws])
        ba = (ctypes.c_void_p * NUM_CONVS)(*[b.data_ptr() for b in bs])
        _check(self.lib.nqa_set_vgg_weights(self.handle, wa, ba))

    def set_alpha_beta(self, alpha: torch.Tensor, beta: torch.Tensor) -> None:
        a = alpha.detach().to(torch.float32).reshape(-1).contiguous()
        b = beta.detach().to(torch.float32).reshape(-1).contiguous()
        assert a.numel() == NUM_CHANNELS and b.numel() == NUM_CHANNELS
        self._set_weights(self.lib.nqa_set_alpha_beta, self.lib.nqa_set_alpha_beta_async, a, b)

    def set_channel_weights(self, a: torch.Tensor, b: torch.Tensor) -> None:
        """Effective per-channel weights, used as given (no division by their sum inside the kernels)."""
        a = a.detach().to(torch.float32).reshape(-1).contiguous()
        b = b.detach().to(torch.float32).reshape(-1).contiguous()
        assert a.numel() == NUM_CHANNELS and b.numel() == NUM_CHANNELS
        self._set_weights(self.lib.nqa_set_channel_weights, self.lib.nqa_set_channel_weights_async, a, b)

    def _set_weights(self, sync_fn, async_fn, a: torch.Tensor, b: torch.Tensor) -> None:
        """Device tensors: stream-ordered copy on torch's current stream (behind the ops that produced them, ahead of
        the scoring call that follows); the caching allocator keeps the sources alive in stream order.  Host tensors:
        the synchronous form."""
        if a.is_cuda and b.is_cuda:
            _check(async_fn(self.handle, _ptr(a), _ptr(b), _stream(self.device)))
        else:
            _check(sync_fn(self.handle, _ptr(a), _ptr(b)))

    # -- scoring ---------------------------------------------------------------------------------
    def dists_forward(self, x: torch.Tensor, y: torch.Tensor, precision: int, want_s1s2: bool = True):
        if x.shape != y.shape:
            raise NerfQAError(f"shape mismatch: x {tuple(x.shape)} vs y {tuple(y.shape)}")
        if x.dim() != 4 or x.shape[1] != 3:
            raise NerfQAError(f"expected [B,3,H,W] frames, got {tuple(x.shape)}")
        x, y = _as_input(x, "x"), _as_input(y, "y")
        if self.ops is not None:
            score, s1s2 = self._op(self.ops.dists_forward, x, y, precision, want_s1s2)
            return score, (s1s2 if want_s1s2 else None)
        b, _, h, w = x.shape
        score = torch.empty(b, device=x.device, dtype=torch.float32)
        s1s2 = torch.empty(b, 2, NUM_CHANNELS, device=x.device, dtype=torch.float32) if want_s1s2 else None
        _check(self.lib.nqa_dists_forward(self.handle, _ptr(x), _ptr(y), b, h, w, precision, _ptr(score),
                                          _ptr(s1s2) if want_s1s2 else None, _stream(x.device)))
        return score, s1s2

    def dists_forward_host(self, x: torch.Tensor, y: torch.Tensor, precision: int, out: Optional[torch.Tensor] = None):
        """x, y: CPU fp32 tensors (pinned for full PCIe rate).  Returns a CPU tensor of scores."""
        if x.is_cuda or y.is_cuda:
            raise NerfQAError("dists_forward_host takes host tensors")
        if x.shape != y.shape or x.dim() != 4 or x.shape[1] != 3:
            raise NerfQAError(f"expected two [B,3,H,W] host tensors, got {tuple(x.shape)} / {tuple(y.shape)}")
        x, y = x.to(torch.float32).contiguous(), y.to(torch.float32).contiguous()
        b, _, h, w = x.shape
        if out is None:
            out = torch.empty(b, dtype=torch.float32).pin_memory()
        _check(self.lib.nqa_dists_forward_host(self.handle, _ptr(x), _ptr(y), b, h, w, precision, _ptr(out)))
        return out

    def _host_pair(self, x: torch.Tensor, y: torch.Tensor, what: str):
        if x.is_cuda or y.is_cuda:
            raise NerfQAError(f"{what} takes host tensors")
        if x.shape != y.shape or x.dim() != 4 or x.shape[1] != 3:
            raise NerfQAError(f"expected two [B,3,H,W] host tensors, got {tuple(x.shape)} / {tuple(y.shape)}")
        return x.to(torch.float32).contiguous(), y.to(torch.float32).contiguous()

    def adists_forward_host(self, x: torch.Tensor, y: torch.Tensor, window: int, precision: int,
                            out: Optional[torch.Tensor] = None):
        """A-DISTS per-frame scores from CPU fp32 frames (pinned for full PCIe rate); returns a CPU tensor."""
        x, y = self._host_pair(x, y, "adists_forward_host")
        b, _, h, w = x.shape
        if out is None:
            out = torch.empty(b, dtype=torch.float32).pin_memory()
        _check(self.lib.nqa_adists_forward_host(self.handle, _ptr(x), _ptr(y), b, h, w, window, precision, _ptr(out)))
        return out

    def dists_adists_forward_host(self, x: torch.Tensor, y: torch.Tensor, window: int, precision: int,
                                  out: Optional[Sequence[torch.Tensor]] = None):
        """(DISTS [B], A-DISTS [B]) from CPU fp32 frames through one trunk evaluation; CPU tensors."""
        x, y = self._host_pair(x, y, "dists_adists_forward_host")
        b, _, h, w = x.shape
        d, a = out if out is not None else (torch.empty(b, dtype=torch.float32).pin_memory(),
                                            torch.empty(b, dtype=torch.float32).pin_memory())
        _check(self.lib.nqa_dists_adists_forward_host(self.handle, _ptr(x), _ptr(y), b, h, w, window, precision,
                                                      _ptr(d), _ptr(a)))
        return d, a

    def features(self, x: torch.Tensor, precision: int):
        if x.dim() != 4 or x.shape[1] != 3:
            raise NerfQAError(f"expected [B,3,H,W] frames, got {tuple(x.shape)}")
        x = _as_input(x, "x")
        if self.ops is not None:
            return list(self._op(self.ops.dists_features, x, precision))
        b, _, h, w = x.shape
        outs = [torch.empty(s, device=x.device, dtype=torch.float32) for s in level_shapes(b, h, w)[1:]]
        arr = (ctypes.c_void_p * (NUM_LEVELS - 1))(*[o.data_ptr() for o in outs])
        _check(self.lib.nqa_dists_features(self.handle, _ptr(x), b, h, w, precision, arr, _stream(x.device)))
        return outs

    def _feat_arrays(self, feats0, feats1):
        if len(feats0) != NUM_LEVELS or len(feats1) != NUM_LEVELS:
            raise NerfQAError("expected six feature levels per image")
        f0 = [_as_input(f, "feats0") for f in feats0]
        f1 = [_as_input(f, "feats1") for f in feats1]
        b, _, h, w = f0[0].shape
        for k, s in enumerate(level_shapes(b, h, w)):
            if tuple(f0[k].shape) != s or tuple(f1[k].shape) != s:
                raise NerfQAError(f"level {k}: expected shape {s}, got {tuple(f0[k].shape)} / {tuple(f1[k].shape)}")
        a0 = (ctypes.c_void_p * NUM_LEVELS)(*[f.data_ptr() for f in f0])
        a1 = (ctypes.c_void_p * NUM_LEVELS)(*[f.data_ptr() for f in f1])
        return f0, f1, a0, a1, b, h, w

    def dists_from_feats(self, feats0, feats1):
        f0, f1, a0, a1, b, h, w = self._feat_arrays(feats0, feats1)
        dev = f0[0].device
        score = torch.empty(b, device=dev, dtype=torch.float32)
        s1s2 = torch.empty(b, 2, NUM_CHANNELS, device=dev, dtype=torch.float32)
        _check(self.lib.nqa_dists_from_feats(self.handle, a0, a1, b, h, w, _ptr(score), _ptr(s1s2), _stream(dev)))
        return score, s1s2

    def dists_backward(self, x: torch.Tensor, y: torch.Tensor, precision: int, grad_score: torch.Tensor,
                       through_trunk: bool):
        """d loss / d x, d loss / d y ([B,3,H,W] fp32) for d loss / d score = grad_score [B]."""
        x, y = _as_input(x, "x"), _as_input(y, "y")
        g = _as_input(grad_score, "grad_score").reshape(-1)
        b, _, h, w = x.shape
        if g.numel() != b:
            raise NerfQAError(f"grad_score must have {b} elements, got {g.numel()}")
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        _check(self.lib.nqa_dists_backward(self.handle, _ptr(x), _ptr(y), b, h, w, precision, _ptr(g),
                                           1 if through_trunk else 0, _ptr(gx), _ptr(gy), _stream(x.device)))
        return gx, gy

    def dists_forward_saved(self, x: torch.Tensor, y: torch.Tensor, precision: int):
        """``(score, s1s2, tape)``: the forward with the trunk activations kept for ``dists_backward_saved``.  ``tape`` is
        None when the library declined to keep them (too large); differentiate with ``dists_backward`` then."""
        x, y = _as_input(x, "x"), _as_input(y, "y")
        if x.shape != y.shape or x.dim() != 4 or x.shape[1] != 3:
            raise NerfQAError(f"expected two [B,3,H,W] batches of the same shape, got {tuple(x.shape)} / {tuple(y.shape)}")
        b, _, h, w = x.shape
        score = torch.empty(b, device=x.device, dtype=torch.float32)
        s1s2 = torch.empty(b, 2, NUM_CHANNELS, device=x.device, dtype=torch.float32)
        tape = ctypes.c_ulonglong(0)
        _check(self.lib.nqa_dists_forward_saved(self.handle, _ptr(x), _ptr(y), b, h, w, precision, _ptr(score), _ptr(s1s2),
                                                ctypes.byref(tape), _stream(x.device)))
        return score, s1s2, (ActivationTape(self, tape.value) if tape.value else None)

    def dists_backward_saved(self, tape: "ActivationTape", x: torch.Tensor, y: torch.Tensor, grad_score: torch.Tensor):
        """Frame gradients from a tape (consumes it)."""
        x, y = _as_input(x, "x"), _as_input(y, "y")
        g = _as_input(grad_score, "grad_score").reshape(-1)
        if g.numel() != x.shape[0]:
            raise NerfQAError(f"grad_score must have {x.shape[0]} elements, got {g.numel()}")
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        tape_id, tape.id = tape.id, 0          # consumed by the call whatever its outcome
        _check(self.lib.nqa_dists_backward_saved(self.handle, tape_id, _ptr(x), _ptr(y), _ptr(g), _ptr(gx), _ptr(gy),
                                                 _stream(x.device)))
        return gx, gy

    def dists_from_feats_backward(self, feats0, feats1, grad_score: torch.Tensor):
        f0, f1, a0, a1, b, h, w = self._feat_arrays(feats0, feats1)
        dev = f0[0].device
        g = _as_input(grad_score, "grad_score").reshape(-1)
        g0 = [torch.empty_like(f) for f in f0]
        g1 = [torch.empty_like(f) for f in f1]
        ga0 = (ctypes.c_void_p * NUM_LEVELS)(*[t.data_ptr() for t in g0])
        ga1 = (ctypes.c_void_p * NUM_LEVELS)(*[t.data_ptr() for t in g1])
        _check(self.lib.nqa_dists_from_feats_backward(self.handle, a0, a1, b, h, w, _ptr(g), ga0, ga1, _stream(dev)))
        return g0, g1

    def adists_forward(self, x: torch.Tensor, y: torch.Tensor, window: int, precision: int) -> torch.Tensor:
        if x.shape != y.shape:
            raise AssertionError(f"shape mismatch: x {tuple(x.shape)} vs y {tuple(y.shape)}")  # ADISTS.py:138
        if x.dim() != 4 or x.shape[1] != 3:
            raise NerfQAError(f"expected [B,3,H,W] frames, got {tuple(x.shape)}")
        x, y = _as_input(x, "x"), _as_input(y, "y")
        if self.ops is not None:
            return self._op(self.ops.adists_forward, x, y, window, precision)
        b, _, h, w = x.shape
        score = torch.empty(b, device=x.device, dtype=torch.float32)
        _check(self.lib.nqa_adists_forward(self.handle, _ptr(x), _ptr(y), b, h, w, window, precision, _ptr(score),
                                           _stream(x.device)))
        return score

    def adists_parts(self, x: torch.Tensor, y: torch.Tensor, window: int, precision: int):
        """(scores [B], channel weights [B,1475], per-level terms [B,6], ps maps: six [B,1,oh_k,ow_k] tensors) -- the
        intermediates of ADISTS.forward for level-by-level parity checks."""
        if x.shape != y.shape:
            raise AssertionError(f"shape mismatch: x {tuple(x.shape)} vs y {tuple(y.shape)}")
        x, y = _as_input(x, "x"), _as_input(y, "y")
        b, _, h, w = x.shape
        dev = x.device
        score = torch.empty(b, device=dev, dtype=torch.float32)
        wts = torch.empty(b, NUM_CHANNELS, device=dev, dtype=torch.float32)
        terms = torch.empty(b, NUM_LEVELS, device=dev, dtype=torch.float32)
        ps = []
        for k, (_, _, hk, wk) in enumerate(level_shapes(b, h, w)):
            n = ctypes.c_size_t()
            _check(self.lib.nqa_adists_ps_elems(h, w, window, k, ctypes.byref(n)))
            shape = (b, 1, hk - window + 1, wk - window + 1) if n.value > 1 or (hk == window and wk == window) else (b, 1, 1, 1)
            assert shape[2] * shape[3] == n.value
            ps.append(torch.empty(shape, device=dev, dtype=torch.float32))
        arr = (ctypes.c_void_p * NUM_LEVELS)(*[p.data_ptr() for p in ps])
        _check(self.lib.nqa_adists_parts(self.handle, _ptr(x), _ptr(y), b, h, w, window, precision, _ptr(score), _ptr(wts),
                                         _ptr(terms), arr, _stream(dev)))
        return score, wts, terms, ps

    def dists_adists_forward(self, x: torch.Tensor, y: torch.Tensor, window: int, precision: int):
        """(DISTS [B], A-DISTS [B]) from one trunk evaluation."""
        if x.shape != y.shape:
            raise NerfQAError(f"shape mismatch: x {tuple(x.shape)} vs y {tuple(y.shape)}")
        x, y = _as_input(x, "x"), _as_input(y, "y")
        if self.ops is not None:
            return self._op(self.ops.dists_adists_forward, x, y, window, precision)
        b, _, h, w = x.shape
        d = torch.empty(b, device=x.device, dtype=torch.float32)
        a = torch.empty(b, device=x.device, dtype=torch.float32)
        _check(self.lib.nqa_dists_adists_forward(self.handle, _ptr(x), _ptr(y), b, h, w, window, precision, _ptr(d), _ptr(a),
                                                 _stream(x.device)))
        return d, a

    def adists_map(self, x: torch.Tensor, y: torch.Tensor, window: int, precision: int):
        """(scores [B], map [B,1,H,W]) of ADISTS.forward(as_map=True)."""
        if x.shape != y.shape:
            raise AssertionError(f"shape mismatch: x {tuple(x.shape)} vs y {tuple(y.shape)}")
        x, y = _as_input(x, "x"), _as_input(y, "y")
        b, _, h, w = x.shape
        score = torch.empty(b, device=x.device, dtype=torch.float32)
        dmap = torch.empty(b, 1, h, w, device=x.device, dtype=torch.float32)
        _check(self.lib.nqa_adists_map(self.handle, _ptr(x), _ptr(y), b, h, w, window, precision, _ptr(score), _ptr(dmap),
                                       _stream(x.device)))
        return score, dmap

    def adists_from_feats(self, feats0, feats1, window: int) -> torch.Tensor:
        f0, f1, a0, a1, b, h, w = self._feat_arrays(feats0, feats1)
        dev = f0[0].device
        score = torch.empty(b, device=dev, dtype=torch.float32)
        _check(self.lib.nqa_adists_from_feats(self.handle, a0, a1, b, h, w, window, _ptr(score), _stream(dev)))
        return score

    def l2pool(self, x: torch.Tensor) -> torch.Tensor:
        x = _as_input(x, "input")
        if self.ops is not None and x.dim() == 4:
            return self._op(self.ops.l2pool, x)
        b, c, h, w = x.shape
        out = torch.empty(b, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1, device=x.device, dtype=torch.float32)
        _check(self.lib.nqa_l2pool(self.handle, _ptr(x), b, c, h, w, _ptr(out), _stream(x.device)))
        return out

    def resize_bilinear(self, x: torch.Tensor, size, antialias: bool = False) -> torch.Tensor:
        """[B,C,H,W] -> [B,C,OH,OW], bilinear, align_corners=False (optionally antialiased like PIL / torchvision)."""
        x = _as_input(x, "x")
        b, c, h, w = x.shape
        oh, ow = int(size[0]), int(size[1])
        out = torch.empty(b, c, oh, ow, device=x.device, dtype=torch.float32)
        _check(self.lib.nqa_resize_bilinear(self.handle, _ptr(x), b, c, h, w, oh, ow, 1 if antialias else 0, _ptr(out),
                                            _stream(x.device)))
        return out

    def conv3x3_relu(self, x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, precision: int) -> torch.Tensor:
        """Op-level conv + bias + ReLU (NCHW fp32 in/out) through the selected kernel; for parity tests."""
        x, w, bias = _as_input(x, "x"), _as_input(w, "w"), _as_input(bias, "bias")
        b, cin, h, wd = x.shape
        cout = w.shape[0]
        out = torch.empty(b, cout, h, wd, device=x.device, dtype=torch.float32)
        _check(self.lib.nqa_conv3x3_relu(self.handle, _ptr(x), _ptr(w), _ptr(bias), b, cin, cout, h, wd, precision,
                                         _ptr(out), _stream(x.device)))
        return out

    # -- introspection ---------------------------------------------------------------------------
    def launch_count(self) -> int:
        n = ctypes.c_uint64()
        _check(self.lib.nqa_launch_count(self.handle, ctypes.byref(n)))
        return int(n.value)

    def graph_replays(self) -> int:
        n = ctypes.c_uint64()
        _check(self.lib.nqa_graph_replays(self.handle, ctypes.byref(n)))
        return int(n.value)

    def profile_enable(self, on: bool) -> None:
        _check(self.lib.nqa_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self):
        """{family: (milliseconds, launches)} accumulated since the previous read (synchronises the device)."""
        ms = (ctypes.c_double * len(KERNEL_FAMILIES))()
        n = (ctypes.c_uint64 * len(KERNEL_FAMILIES))()
        _check(self.lib.nqa_profile_read(self.handle, ms, n))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_FAMILIES)}

    def measure_rate(self, which: str) -> float:
        """Measured multiply-adds per second: "fp32_fma" (packed FFMA2 on the CUDA cores) or "mma_sync_bf16"."""
        v = ctypes.c_double()
        _check(self.lib.nqa_measure_rate(self.handle, {"fp32_fma": 0, "mma_sync_bf16": 1}[which], ctypes.byref(v)))
        return float(v.value)

    def workspace_bytes(self) -> int:
        n = ctypes.c_size_t()
        _check(self.lib.nqa_workspace_bytes(self.handle, ctypes.byref(n)))
        return int(n.value)

    def micro_batch(self, b: int, h: int, w: int, precision: int) -> int:
        n = ctypes.c_int()
        _check(self.lib.nqa_micro_batch(self.handle, b, h, w, precision, ctypes.byref(n)))
        return int(n.value)
