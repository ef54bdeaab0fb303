"""Frame preparation in front of the scorers, on the GPU.

The reference's offline scorers resize every decoded frame on the host before `model(ref, render)`; the four variants of
test2_prep.py write the `DISTS` / `DISTS_square` / `DISTS_full_size` / `DISTS_pixel_count` columns of scores_aspect*.csv.
Here full-resolution frames are uploaded once and resized where they are scored (``nqa_resize_bilinear``):

  ``aspect``       short side -> 256, aspect kept, antialiased   test2_prep.py:109-114, DISTS_pt.py:210-217
                   (transforms.functional.resize(image, 256) on the PIL image; only if min(size) > 256)
  ``square``       256 x 256, antialiased                         test2_prep.py:221-227  (only if min(size) > 256)
  ``full_size``    no resize                                      test2_prep.py:324-327
  ``pixel_count``  aspect kept, H * W ~ 256 * 256, plain bilinear test2_prep.py:424-439  (always applied)
  ``interp256``    256 x 256, plain bilinear                      prep.py:91-96 (load_video_frames)

CUDA tensors only -- like the scorers there is no CPU fallback."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from . import _lib

MODES = ("aspect", "square", "full_size", "pixel_count", "interp256")
_ctx: Dict[int, _lib.Context] = {}


def target_size(mode: str, h: int, w: int) -> Tuple[int, int, bool]:
    """(OH, OW, antialias) the reference's loader produces for an H x W frame."""
    if mode == "full_size":
        return h, w, False
    if mode == "aspect":
        if min(h, w) <= 256:
            return h, w, False
        # torchvision resize with an int size: short side = size, long side = int(size * long / short)
        return (256, int(256 * w / h), True) if h <= w else (int(256 * h / w), 256, True)
    if mode == "square":
        return (256, 256, True) if min(h, w) > 256 else (h, w, False)
    if mode == "pixel_count":
        if w >= h:
            ratio = float(w) / float(h)
            oh = math.sqrt(256 * 256 / ratio)
            ow = int(ratio * oh)
        else:
            ratio = float(h) / float(w)
            ow = math.sqrt(256 * 256 / ratio)
            oh = int(ratio * ow)
        return int(oh), int(ow), False      # F.interpolate truncates a float size like int()
    if mode == "interp256":
        return 256, 256, False
    raise ValueError(f"mode must be one of {MODES}, got {mode!r}")


def resize(x: torch.Tensor, size: Tuple[int, int], antialias: bool = False) -> torch.Tensor:
    """Bilinear resize of [B,C,H,W] CUDA frames (align_corners=False), optionally antialiased."""
    if not x.is_cuda:
        raise _lib.NerfQAError("frames.resize: CUDA tensor required (no CPU fallback)")
    idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
    if idx not in _ctx:
        _ctx[idx] = _lib.Context(torch.device("cuda", idx))
    return _ctx[idx].resize_bilinear(x, size, antialias)


def prepare_frames(x: torch.Tensor, mode: str = "aspect") -> torch.Tensor:
    """The reference loader's resize for one of its variants, on [B,3,H,W] fp32 frames in [0,1]."""
    oh, ow, aa = target_size(mode, x.shape[2], x.shape[3])
    if (oh, ow) == (x.shape[2], x.shape[3]):
        return x
    return resize(x, (oh, ow), aa)
