"""Drop-in ``ADISTS`` scorer (``nerf_qa/ADISTS/ADISTS.py``, paths relative to /root/reference): same
constructor, ``forward(x, y, as_loss=True, as_map=False)`` and state_dict layout, computed by the C-ABI
library.  CUDA tensors only — there is no eager fallback."""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .dists import _TrunkModule, _refuse_grad, build_vgg_stages, try_load_pretrained_vgg, L2pooling


class Downsample(L2pooling):
    """Same operator under the name A-DISTS uses (ADISTS.py:16-31)."""


class ADISTS(_TrunkModule):
    def __init__(self, window_size: int = 21, precision: str = "bf16"):
        super().__init__()
        build_vgg_stages(self, Downsample)
        if not try_load_pretrained_vgg(self):
            import warnings
            warnings.warn("torchvision VGG16 ImageNet weights not found in the local torch hub cache; the trunk keeps "
                          "its random init until a state_dict is loaded (the reference downloads them).")
        for param in self.parameters():
            param.requires_grad = False

        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, -1, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, -1, 1, 1))

        self.chns = [3, 64, 128, 256, 512, 512]
        self.windows = nn.ParameterList()
        self.window_size = window_size
        for k in range(len(self.chns)):
            self.windows.append(self.create_window(self.window_size, self.window_size / 3, self.chns[k]))
        self._init_runtime(precision)

    # kept for state_dict / attribute compatibility (ADISTS.py:102-110); the kernels rebuild the separable taps
    def gaussian(self, window_size, sigma):
        gauss = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
        return gauss / gauss.sum()

    def create_window(self, window_size, window_sigma, channel):
        _1d = self.gaussian(window_size, window_sigma).unsqueeze(1)
        _2d = _1d.mm(_1d.t()).float().unsqueeze(0).unsqueeze(0)
        window = _2d.expand(channel, 1, window_size, window_size).contiguous()
        return nn.Parameter(window, requires_grad=False)

    def forward(self, x, y, as_loss: bool = True, as_map: bool = False):
        """``as_loss=False`` -> per-frame scores ``[B]``; ``as_loss=True`` (the reference default) -> the scalar
        ``1 - mean_b D`` (ADISTS.py:192-197).  Asymmetric: probability maps and entropy weights come from ``x``."""
        assert x.shape == y.shape                                  # ADISTS.py:138
        # the reference's default (as_loss=True) runs the trunk under autograd (ADISTS.py:139-141); this path has no
        # A-DISTS backward, so a call that expects a gradient raises instead of returning a graph-less scalar
        _refuse_grad("ADISTS.forward", x, y)
        ctx = self._context(x.device)
        if as_map:
            # ADISTS.py:163,189: D_map_full is [B,H,W] and each level adds a [B,1,H,W] tensor, so for B > 1 the
            # reference's result broadcasts to [B,B,H,W] with the per-sample map repeated along dim 1 (upstream quirk)
            _, dmap = ctx.adists_map(x, y, self.window_size, self._prec())
            return dmap if x.shape[0] == 1 else dmap.expand(-1, x.shape[0], -1, -1)
        score = ctx.adists_forward(x, y, self.window_size, self._prec())
        if as_loss:
            return score.mean()        # 1 - mean(D) == mean(1 - D)
        return score

    def score_host(self, x: torch.Tensor, y: torch.Tensor, device=None) -> torch.Tensor:
        """Per-frame scores from HOST frames (pinned recommended), as a CPU tensor: what
        ``adists_model(ref.to(device), render.to(device), as_loss=False).cpu()`` does in prep.py:186-187, with the frames
        streamed to the GPU in micro-batches overlapped with compute."""
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        ctx = self._context(dev)
        return ctx.adists_forward_host(x, y, self.window_size, self._prec())

    def forward_from_feats(self, feats_x: Sequence[torch.Tensor], feats_y: Sequence[torch.Tensor]) -> torch.Tensor:
        """The statistics half of forward on six NCHW fp32 levels (what ``forward_once`` returns)."""
        ctx = self._context(feats_x[0].device)
        return ctx.adists_from_feats(feats_x, feats_y, self.window_size)


def prepare_image(image, resize: bool = True) -> torch.Tensor:
    """ADISTS.py:200-204."""
    from torchvision import transforms
    if resize and min(image.size) > 256:
        image = transforms.functional.resize(image, 256)
    image = transforms.ToTensor()(image)
    return image.unsqueeze(0)
