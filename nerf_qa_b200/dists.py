"""Drop-in ``DISTS`` scorer: same constructor, methods, attributes and state_dict layout as
``nerf_qa/DISTS_pytorch/DISTS_pt.py`` (reference paths relative to /root/reference), with every forward
routed through the C-ABI library (hand-written sm_100a kernels).  No eager-PyTorch fallback exists: CPU
tensors raise.

Differences a caller can observe are listed in DESIGN.md ("API deviations"): ``precision=`` selects the
trunk arithmetic.  ``DISTSOriginal`` / ``DISTSSoftmax`` are the reference's ``DISTS_pt_original`` / ``DISTS_pt_softmax``
weight-normalisation flavours (same kernels, different host-side weight transform).
"""
from __future__ import annotations

import os
import warnings
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .synth import CHNS, CONV_STAGE, VGG_CONVS, VGG_FEATURE_IDX

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class L2pooling(nn.Module):
    """``sqrt(hanning3x3(x^2, stride 2, pad 1) + 1e-12)`` (DISTS_pt.py:11-25).  ``filter_size=5`` yields a
    3x3 kernel because the reference trims ``np.hanning(5)`` to its interior."""

    def __init__(self, filter_size: int = 5, stride: int = 2, channels: Optional[int] = None, pad_off: int = 0):
        super().__init__()
        if filter_size != 5 or stride != 2:
            raise NotImplementedError("nerf_qa_b200 implements the reference's L2pooling(filter_size=5, stride=2)")
        self.padding = (filter_size - 2) // 2
        self.stride = stride
        self.channels = channels
        a = np.hanning(filter_size)[1:-1]
        g = torch.Tensor(a[:, None] * a[None, :])
        g = g / torch.sum(g)
        self.register_buffer("filter", g[None, None, :, :].repeat((self.channels, 1, 1, 1)))
        self._ctx: Dict[int, _lib.Context] = {}

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if not input.is_cuda:
            raise _lib.NerfQAError("L2pooling: CUDA tensor required (no CPU fallback)")
        _refuse_grad("L2pooling.forward", input)
        idx = input.device.index
        if idx not in self._ctx:
            self._ctx[idx] = _lib.Context(input.device)
        return self._ctx[idx].l2pool(input)


def _refuse_grad(what: str, *tensors: torch.Tensor) -> None:
    """The reference's ops are differentiable everywhere; here only DISTS.forward / forward_from_feats are.  A call
    that would silently drop out of the autograd graph raises instead."""
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
        raise NotImplementedError(
            f"{what}: gradients are not implemented on this path (nerf_qa_b200 differentiates DISTS.forward and "
            "DISTS.forward_from_feats); call it under torch.no_grad() or detach the inputs")


def build_vgg_stages(module: nn.Module, pool_cls) -> None:
    """stage1..stage5 with the reference's child names (DISTS_pt.py:31-49) so state_dict keys such as
    ``stage3.10.weight`` / ``stage2.4.filter`` line up with reference checkpoints and torchvision's VGG16."""
    stages = [nn.Sequential() for _ in range(5)]
    pool_idx = {2: 4, 3: 9, 4: 16, 5: 23}
    for st in range(2, 6):
        stages[st - 1].add_module(str(pool_idx[st]), pool_cls(channels=CHNS[st - 1]))
    for i, idx in enumerate(VGG_FEATURE_IDX):
        cin, cout = VGG_CONVS[i]
        st = CONV_STAGE[i]
        stages[st - 1].add_module(str(idx), nn.Conv2d(cin, cout, kernel_size=3, padding=1))
        stages[st - 1].add_module(str(idx + 1), nn.ReLU(inplace=True))
    for k, s in enumerate(stages):
        setattr(module, f"stage{k + 1}", s)


def try_load_pretrained_vgg(module: nn.Module) -> bool:
    """The reference downloads torchvision's ImageNet VGG16 (DISTS_pt.py:30).  Offline we only look in the
    local torch hub cache; otherwise the (random) init stays and the caller is expected to load weights."""
    path = os.path.join(torch.hub.get_dir(), "checkpoints", "vgg16-397923af.pth")
    if not os.path.exists(path):
        return False
    sd = torch.load(path, map_location="cpu")
    own = module.state_dict()
    for i, idx in enumerate(VGG_FEATURE_IDX):
        st = CONV_STAGE[i]
        own[f"stage{st}.{idx}.weight"] = sd[f"features.{idx}.weight"]
        own[f"stage{st}.{idx}.bias"] = sd[f"features.{idx}.bias"]
    module.load_state_dict(own)
    return True


def conv_parameters(module: nn.Module):
    ws, bs = [], []
    for i, idx in enumerate(VGG_FEATURE_IDX):
        conv = getattr(module, f"stage{CONV_STAGE[i]}")._modules[str(idx)]
        ws.append(conv.weight)
        bs.append(conv.bias)
    return ws, bs


class _TrunkModule(nn.Module):
    """Shared plumbing of DISTS and ADISTS: per-device library contexts and weight upload tracking."""

    precision: str

    def _init_runtime(self, precision: str) -> None:
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.precision = precision
        self._contexts: Dict[int, _lib.Context] = {}

    def _prec(self) -> int:
        return _lib.PRECISIONS[self.precision]

    def _context(self, device: torch.device) -> _lib.Context:
        if device.type != "cuda":
            raise _lib.NerfQAError("nerf_qa_b200 runs on CUDA (B200) tensors only; there is no CPU fallback")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        ctx = self._contexts.get(idx)
        if ctx is None:
            ctx = _lib.Context(torch.device("cuda", idx))
            self._contexts[idx] = ctx
        ws, bs = conv_parameters(self)
        # (address, version counter) of the 26 tensors: catches .to(), load_state_dict, optimizer steps and in-place ops on
        # the parameters.  In-place writes through `.data` (p.data.copy_(...)) bump no counter: call invalidate_weights().
        key = (self._weights_epoch,) + tuple((p.data_ptr(), p._version) for p in ws + bs)
        if ctx.weights_key != key:
            ctx.set_vgg_weights(ws, bs)
            ctx.weights_key = key
        return ctx

    _weights_epoch = 0

    def invalidate_weights(self) -> None:
        """Forces the next call to re-upload (and re-pack) the VGG weights.  Needed only after writing conv weights
        through ``.data`` in place, which torch's version counters do not see; ``load_state_dict``, ``.to()`` and ordinary
        in-place ops are detected (and call this too)."""
        self._weights_epoch += 1

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_weights()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_weights()
        return out

    def forward_once(self, x: torch.Tensor) -> List[torch.Tensor]:
        """[x, relu1_2, relu2_2, relu3_3, relu4_3, relu5_3] as NCHW fp32 (DISTS_pt.py:91-103, ADISTS.py:112-125).
        Level 0 is the RAW input, not the normalised tensor."""
        _refuse_grad("forward_once", x)
        ctx = self._context(x.device)
        return [x] + ctx.features(x, self._prec())

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_contexts"] = {}
        return state


class _WeightedScore(torch.autograd.Function):
    """score_b = 1 - sum_c (a_c S1_bc + b_c S2_bc) with EFFECTIVE channel weights a, b used as given
    (nqa_set_channel_weights): the weight normalisation (alpha / (sum alpha + sum beta) in DISTS_pt.py:127-129; relu /
    detach / softmax in the `_original` / `_softmax` flavours) stays in torch and is differentiated by autograd.
    Backward: frames through nqa_dists_backward (the whole trunk when require_grad=True, level-0 statistics only
    otherwise -- the reference's forward_once returns the raw x as level 0 even under no_grad); weights in closed form
    from the saved per-channel S1/S2: d score_b / d a_c = -S1_bc, d score_b / d b_c = -S2_bc."""

    @staticmethod
    def forward(ctx, x, y, a, b, lib_ctx, prec, through_trunk):
        lib_ctx.set_channel_weights(a, b)
        ctx.tape = None
        if through_trunk and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            # like torch autograd: keep the trunk activations for the backward (the library declines -> tape None ->
            # activation recomputation in nqa_dists_backward)
            score, s1s2, ctx.tape = lib_ctx.dists_forward_saved(x, y, prec)
        else:
            score, s1s2 = lib_ctx.dists_forward(x, y, prec, want_s1s2=True)
        ctx.save_for_backward(x, y, s1s2, a, b)
        ctx.lib_ctx, ctx.prec, ctx.through_trunk = lib_ctx, prec, through_trunk
        return score

    @staticmethod
    def backward(ctx, grad_out):
        x, y, s1s2, a, b = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        gx = gy = ga = gb = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.lib_ctx.set_channel_weights(a, b)
            if ctx.tape is not None and ctx.tape.id:
                gx, gy = ctx.lib_ctx.dists_backward_saved(ctx.tape, x, y, grad_out)
            else:                               # no tape, or a second backward through a retained graph
                gx, gy = ctx.lib_ctx.dists_backward(x, y, ctx.prec, grad_out, ctx.through_trunk)
            gx = gx if ctx.needs_input_grad[0] else None
            gy = gy if ctx.needs_input_grad[1] else None
        if ctx.needs_input_grad[2]:
            ga = -(grad_out.unsqueeze(1) * s1s2[:, 0]).sum(0).reshape(a.shape)
        if ctx.needs_input_grad[3]:
            gb = -(grad_out.unsqueeze(1) * s1s2[:, 1]).sum(0).reshape(b.shape)
        return gx, gy, ga, gb, None, None, None


class _WeightedFromFeats(torch.autograd.Function):
    """forward_from_feats under autograd (DISTS_pt.py:181-208): gradients w.r.t. the 2 x 6 feature tensors
    (what the NR models back-propagate, model_nr_v8.py:258-265) and the effective channel weights."""

    @staticmethod
    def forward(ctx, a, b, lib_ctx, *feats):
        lib_ctx.set_channel_weights(a, b)
        score, s1s2 = lib_ctx.dists_from_feats(feats[:6], feats[6:])
        ctx.save_for_backward(s1s2, a, b, *feats)
        ctx.lib_ctx = lib_ctx
        return score

    @staticmethod
    def backward(ctx, grad_out):
        s1s2, a, b = ctx.saved_tensors[:3]
        feats = ctx.saved_tensors[3:]
        grad_out = grad_out.contiguous()
        ga = gb = None
        gf = [None] * 12
        if any(ctx.needs_input_grad[3:]):
            ctx.lib_ctx.set_channel_weights(a, b)
            g0, g1 = ctx.lib_ctx.dists_from_feats_backward(feats[:6], feats[6:], grad_out)
            gf = [g if need else None for g, need in zip(list(g0) + list(g1), ctx.needs_input_grad[3:])]
        if ctx.needs_input_grad[0]:
            ga = -(grad_out.unsqueeze(1) * s1s2[:, 0]).sum(0).reshape(a.shape)
        if ctx.needs_input_grad[1]:
            gb = -(grad_out.unsqueeze(1) * s1s2[:, 1]).sum(0).reshape(b.shape)
        return (ga, gb, None, *gf)


class DISTS(_TrunkModule):
    def __init__(self, load_weights: bool = True, from_feats: bool = False, precision: str = "bf16"):
        super().__init__()
        build_vgg_stages(self, L2pooling)
        if not try_load_pretrained_vgg(self):
            warnings.warn("torchvision VGG16 ImageNet weights not found in the local torch hub cache; the trunk keeps "
                          "its random init until a state_dict is loaded (the reference downloads them).")
        for param in self.parameters():
            param.requires_grad = False

        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, -1, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, -1, 1, 1))

        self.chns = [3, 64, 128, 256, 512, 512]
        self.register_parameter("alpha", nn.Parameter(torch.randn(1, sum(self.chns), 1, 1)))
        self.register_parameter("beta", nn.Parameter(torch.randn(1, sum(self.chns), 1, 1)))
        self.alpha.data.normal_(0.1, 0.01)
        self.beta.data.normal_(0.1, 0.01)
        if load_weights:
            # the reference reads sys.prefix/weights.pt (DISTS_pt.py:63); the same learned alpha/beta ship here
            ab = np.load(os.path.join(_DATA, "dists_alpha_beta.npy"))
            self.alpha.data = torch.from_numpy(ab[0].copy()).view(1, -1, 1, 1)
            self.beta.data = torch.from_numpy(ab[1].copy()).view(1, -1, 1, 1)
        self._init_runtime(precision)

    # -- weights ---------------------------------------------------------------------------------
    def project_weights(self) -> None:
        """DISTS_pt.py:82-89."""
        lower_bound = torch.zeros_like(self.alpha.data)
        lower_bound[:, :3, :, :] = 0.02
        alpha = torch.max(self.alpha.data, lower_bound)
        beta = torch.max(self.beta.data, lower_bound)
        weight_sum = torch.cat([alpha, beta], dim=1).sum()
        self.alpha.data = alpha / weight_sum
        self.beta.data = beta / weight_sum

    def _effective_weights(self):
        """The per-channel weights the score is formed with (DISTS_pt.py:127-129): alpha, beta over their joint sum.
        Differentiable torch ops on 2 x 1475 values; the flavours below override this."""
        w_sum = self.alpha.sum() + self.beta.sum()
        return self.alpha / w_sum, self.beta / w_sum

    _squeeze_score = False      # DISTS_pt_original returns score.squeeze() (0-d for B = 1)

    def _sync_alpha_beta(self, ctx: _lib.Context) -> None:
        # every call, like the reference reads the live parameters (2 x 5.9 KB, stream-ordered): `.data` edits such as
        # project_weights() or a checkpoint copy are never missed
        ctx.set_alpha_beta(self.alpha, self.beta)

    # -- scoring ---------------------------------------------------------------------------------
    def forward(self, x, y, require_grad: bool = False, batch_average: bool = False, warp=None, certainty=None):
        """Per-frame DISTS scores ``[B]`` (or their mean).  ``warp`` / ``certainty`` are accepted and ignored
        exactly like the reference (DISTS_pt.py:105)."""
        ctx = self._context(x.device)
        wants_grad = torch.is_grad_enabled() and (x.requires_grad or y.requires_grad or self.alpha.requires_grad
                                                  or self.beta.requires_grad)
        if wants_grad:
            # require_grad=False keeps the trunk out of the graph like the reference's torch.no_grad() block; the
            # level-0 statistics of the raw frames stay differentiable (DISTS_pt.py:103, 109-112)
            a, b = self._effective_weights()
            score = _WeightedScore.apply(x, y, a, b, ctx, self._prec(), bool(require_grad))
        else:
            self._sync_weights_nograd(ctx)
            score, _ = ctx.dists_forward(x, y, self._prec(), want_s1s2=False)
        if self._squeeze_score:
            score = score.squeeze()
        return score.mean() if batch_average else score

    def _sync_weights_nograd(self, ctx: _lib.Context) -> None:
        """Inference: the plain flavour hands raw alpha/beta to the library (it divides by their joint sum in fp64);
        flavours with their own normalisation upload the effective weights."""
        if type(self)._effective_weights is DISTS._effective_weights:
            self._sync_alpha_beta(ctx)
        else:
            with torch.no_grad():
                a, b = self._effective_weights()
            ctx.set_channel_weights(a, b)

    def forward_from_feats(self, feats0: Sequence[torch.Tensor], feats1: Sequence[torch.Tensor],
                           batch_average: bool = False):
        """DISTS_pt.py:181-208."""
        ctx = self._context(feats0[0].device)
        feats = list(feats0) + list(feats1)
        if torch.is_grad_enabled() and (self.alpha.requires_grad or self.beta.requires_grad
                                        or any(f.requires_grad for f in feats)):
            a, b = self._effective_weights()
            score = _WeightedFromFeats.apply(a, b, ctx, *feats)
        else:
            self._sync_weights_nograd(ctx)
            score, _ = ctx.dists_from_feats(feats0, feats1)
        return score.mean() if batch_average else score

    def score_with_adists(self, x: torch.Tensor, y: torch.Tensor, window_size: int = 21):
        """(DISTS [B], A-DISTS [B]) of the batch from ONE trunk evaluation -- the pair of calls prep.py:186-190 /
        test2_prep.py:151-155 make per batch (``adists_model(ref, render, as_loss=False)``, ``dists_model(ref,
        render)``), whose trunks are the same VGG16.  DISTS equals ``forward`` up to the summation order of the fp32 moment
        partials (1e-6 relative), A-DISTS equals ``ADISTS.forward`` up to the rounding of the pooled activations."""
        ctx = self._context(x.device)
        self._sync_weights_nograd(ctx)
        return ctx.dists_adists_forward(x, y, window_size, self._prec())

    def score_host(self, x: torch.Tensor, y: torch.Tensor, device: Optional[torch.device] = None) -> torch.Tensor:
        """End-to-end call from HOST tensors (pinned recommended): streams frames to the GPU in micro-batches
        overlapped with compute and returns CPU scores — what ``model(ref.to(device), dist.to(device)).cpu()``
        does in prep.py:186-190, without staging the whole batch on the device first."""
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        ctx = self._context(dev)
        self._sync_weights_nograd(ctx)
        return ctx.dists_forward_host(x, y, self._prec())

    def score_with_adists_host(self, x: torch.Tensor, y: torch.Tensor, window_size: int = 21,
                               device: Optional[torch.device] = None):
        """``score_with_adists`` from HOST frames: (DISTS [B], A-DISTS [B]) as CPU tensors, frames streamed to the GPU
        in micro-batches overlapped with compute (the whole per-batch body of prep.py:181-190)."""
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        ctx = self._context(dev)
        self._sync_weights_nograd(ctx)
        return ctx.dists_adists_forward_host(x, y, window_size, self._prec())


class _FlavourConfig:
    """The four knobs the reference's flavours read from the global ``wandb.config`` (DISTS_pt_original.py:69-70,
    111-118; DISTS_pt_softmax.py:124).  Explicit constructor arguments win; otherwise a live ``wandb.config`` is
    consulted when wandb is importable; otherwise the reference's argparse defaults (run_final.py:64-70)."""
    DEFAULTS = {"weight_lower_bound": 1e-4, "alpha_beta_ratio": 2.0, "dists_weight_norm": "off", "detach_beta": "False"}

    def __init__(self, **overrides):
        self.values = {k: v for k, v in overrides.items() if v is not None}

    def get(self, key):
        if key in self.values:
            return self.values[key]
        try:
            import wandb
            cfg = wandb.config
            if cfg is not None and key in cfg:
                return cfg[key]
        except Exception:
            pass
        return self.DEFAULTS[key]


class DISTSOriginal(DISTS):
    """``nerf_qa/DISTS_pytorch/DISTS_pt_original.py`` (the FR-training flavour, model_stats.py / model.py): alpha/beta
    clamped from below at load (:65-72), optional ``relu`` / ``w_sum_detach`` weight normalisation and ``detach_beta``
    (:111-121), ``score.squeeze()`` (:134), ``project_weights`` with a uniform lower bound (:88-95)."""

    _squeeze_score = True

    def __init__(self, load_weights: bool = True, from_feats: bool = False, precision: str = "bf16",
                 weight_lower_bound=None, alpha_beta_ratio=None, dists_weight_norm=None, detach_beta=None):
        super().__init__(load_weights=load_weights, from_feats=from_feats, precision=precision)
        self.config = _FlavourConfig(weight_lower_bound=weight_lower_bound, alpha_beta_ratio=alpha_beta_ratio,
                                     dists_weight_norm=dists_weight_norm, detach_beta=detach_beta)
        if load_weights:
            self.original_alpha = self.alpha.data.clone()
            self.original_beta = self.beta.data.clone()
            lb, ratio = float(self.config.get("weight_lower_bound")), float(self.config.get("alpha_beta_ratio"))
            self.alpha.data = torch.clamp(self.alpha.data, min=lb * ratio)
            self.beta.data = torch.clamp(self.beta.data, min=lb)

    def project_weights(self) -> None:
        lower_bound = torch.zeros_like(self.alpha.data) + float(self.config.get("weight_lower_bound"))
        alpha = torch.max(self.alpha.data, lower_bound * float(self.config.get("alpha_beta_ratio")))
        beta = torch.max(self.beta.data, lower_bound)
        weight_sum = torch.cat([alpha, beta], dim=1).sum()
        self.alpha.data = alpha / weight_sum
        self.beta.data = beta / weight_sum

    def _effective_weights(self):
        modes = str(self.config.get("dists_weight_norm")).split("+")
        alpha = torch.relu(self.alpha) if "relu" in modes else self.alpha
        beta = torch.relu(self.beta) if "relu" in modes else self.beta
        if str(self.config.get("detach_beta")) == "True":
            beta = beta.detach()
        w_sum = alpha.sum() + beta.sum()
        if "w_sum_detach" in modes:
            w_sum = w_sum.detach()
        return alpha / w_sum, beta / w_sum


class DISTSSoftmax(DISTS):
    """``nerf_qa/DISTS_pytorch/DISTS_pt_softmax.py``: alpha/beta are logits (initialised to log(max(w, 0) + 1e-10) of
    the shipped weights, :65-78); the channel weights are the softmax over all 2950 logits (:117-123)."""

    def __init__(self, load_weights: bool = True, from_feats: bool = False, precision: str = "bf16", detach_beta=None):
        super().__init__(load_weights=load_weights, from_feats=from_feats, precision=precision)
        self.config = _FlavourConfig(detach_beta=detach_beta)
        if load_weights:
            self.original_alpha = self.alpha.data.clone()
            self.original_beta = self.beta.data.clone()
            w = torch.cat([self.alpha.data, self.beta.data], dim=1)
            logits = torch.log(torch.clamp(w, min=0.0) + 1e-10)
            a_logits, b_logits = torch.split(logits, [self.alpha.shape[1], self.beta.shape[1]], dim=1)
            self.alpha.data = a_logits.contiguous()
            self.beta.data = b_logits.contiguous()

    def _effective_weights(self):
        w = torch.softmax(torch.cat([self.alpha, self.beta], dim=1), dim=1)
        a, b = torch.split(w, self.alpha.shape[1], dim=1)
        if str(self.config.get("detach_beta")) == "True":
            b = b.detach()
        return a, b


def prepare_image(image, resize: bool = True, keep_aspect_ratio: bool = False) -> torch.Tensor:
    """PIL image -> [1,3,H,W] fp32 in [0,1], short side limited to 256 (DISTS_pt.py:210-217)."""
    from torchvision import transforms
    if resize and min(image.size) > 256:
        if keep_aspect_ratio:
            image = transforms.functional.resize(image, 256)
        else:
            image = transforms.functional.resize(image, (256, 256))
    image = transforms.ToTensor()(image)
    return image.unsqueeze(0)
