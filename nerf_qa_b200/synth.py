"""Deterministic synthetic frames and VGG16 weights (no network, no datasets).

The reference builds its trunk from ``torchvision.models.vgg16(pretrained=True)``
(/root/reference/nerf_qa/DISTS_pytorch/DISTS_pt.py:30) and scores decoded video frames
(/root/reference/prep.py:80-100).  Neither the ImageNet weights nor the datasets exist offline, so
benchmarks and parity tests use the generators below: smooth "NeRF-like" content with
five distortion types, and a seeded random-init VGG16.  Everything is generated on the CPU with an
explicit ``torch.Generator`` so the same tensors appear in this container and on the GPU box.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.nn.functional as F

# torchvision vgg16().features conv layout: (cin, cout) of the 13 convs in order; the scorer's
# stages hold them as 2,2,3,3,3 (DISTS_pt.py:36-49).
VGG_CONVS: Tuple[Tuple[int, int], ...] = (
    (3, 64), (64, 64),
    (64, 128), (128, 128),
    (128, 256), (256, 256), (256, 256),
    (256, 512), (512, 512), (512, 512),
    (512, 512), (512, 512), (512, 512),
)
STAGE_CONVS: Tuple[int, ...] = (2, 2, 3, 3, 3)
CHNS: Tuple[int, ...] = (3, 64, 128, 256, 512, 512)
# index of each conv inside torchvision's vgg16().features (== state_dict key stage{k}.{idx}.weight)
VGG_FEATURE_IDX: Tuple[int, ...] = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)
CONV_STAGE: Tuple[int, ...] = (1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5)

DISTORTIONS = ("noise05", "noise20", "blur7", "roll5", "independent")


def make_vgg_weights(seed: int = 1234) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """13 (weight[OIHW fp32], bias[O fp32]) pairs: kaiming-normal fan_out like torchvision's init,
    plus small non-zero biases so the bias path of the conv epilogue is exercised."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    ws, bs = [], []
    for cin, cout in VGG_CONVS:
        std = math.sqrt(2.0 / (cout * 9))
        ws.append(torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float32) * std)
        bs.append(torch.randn(cout, generator=g, dtype=torch.float32) * 0.05)
    return ws, bs


def _box_blur(img: torch.Tensor, k: int) -> torch.Tensor:
    c = img.shape[1]
    ker = torch.full((c, 1, k, k), 1.0 / (k * k), dtype=img.dtype)
    return F.conv2d(F.pad(img, (k // 2,) * 4, mode="reflect"), ker, groups=c)


def make_reference_frames(b: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """[b,3,h,w] fp32 in [0,1]: uniform noise, 9x9 reflect box blur, per-image min-max rescale."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.rand(b, 3, h, w, generator=g, dtype=torch.float32)
    x = _box_blur(x, 9)
    lo = x.amin(dim=(1, 2, 3), keepdim=True)
    hi = x.amax(dim=(1, 2, 3), keepdim=True)
    return ((x - lo) / (hi - lo)).contiguous()


def distort(ref: torch.Tensor, kind: str, seed: int = 1) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    if kind == "noise05":
        out = ref + 0.05 * torch.randn(ref.shape, generator=g)
    elif kind == "noise20":
        out = ref + 0.2 * torch.randn(ref.shape, generator=g)
    elif kind == "blur7":
        out = _box_blur(ref, 7)
    elif kind == "roll5":
        out = torch.roll(ref, shifts=5, dims=3)
    elif kind == "independent":
        out = make_reference_frames(ref.shape[0], ref.shape[2], ref.shape[3], seed=seed + 7919)
    else:
        raise ValueError(f"unknown distortion {kind!r}")
    return out.clamp(0.0, 1.0).contiguous()


def make_frame_pairs(b: int, h: int, w: int, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(ref, dist) batches; the distortion type cycles over the batch index."""
    ref = make_reference_frames(b, h, w, seed=seed)
    dist = torch.empty_like(ref)
    for i in range(b):
        dist[i : i + 1] = distort(ref[i : i + 1], DISTORTIONS[i % len(DISTORTIONS)], seed=seed * 131 + i + 1)
    return ref, dist
