"""CPU ORACLE — test infrastructure only, never a product path.

A functional torch-CPU restatement of the DISTS / A-DISTS scorers of kobejean/nerf-qa.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product (``nerf_qa_b200``) must never call it and fails loudly when its CUDA
library is missing.

Parity status: PINNED against the reference's own modules.  The reference owns no executable tests or
golden vectors for this path, so ``tests/golden/make_golden.py`` imports the
unmodified reference classes from /root/reference (two offline shims: random-init VGG16 instead of the
ImageNet download, alpha/beta from the shipped weights.pt), runs them on seeded synthetic inputs and
commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` holds this restatement to
those vectors.

Arithmetic follows the reference in fp32 (or fp64 when the inputs are double) with the same ATen ops the
reference's CPU path dispatches to, so it also serves as the timed host-CPU baseline ("port").  The functions follow
the device of their inputs (constants are created there), so a test can also run them as eager PyTorch on a GPU -- the
reference's own GPU mode -- for context timings.

File:line citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

CHNS = (3, 64, 128, 256, 512, 512)          # nerf_qa/DISTS_pytorch/DISTS_pt.py:57
STAGE_CONVS = (2, 2, 3, 3, 3)               # DISTS_pt.py:36-49 (vgg16.features idx 0-29)
IMAGENET_MEAN = (0.485, 0.456, 0.406)       # DISTS_pt.py:54
IMAGENET_STD = (0.229, 0.224, 0.225)        # DISTS_pt.py:55


# ----------------------------------------------------------------------------------------------
# trunk
# ----------------------------------------------------------------------------------------------
def hanning_filter(dtype=torch.float32) -> torch.Tensor:
    """3x3 blur of L2pooling: np.hanning(5)[1:-1] outer product, normalised (DISTS_pt.py:16-19)."""
    a = np.hanning(5)[1:-1]
    g = torch.tensor(a[:, None] * a[None, :], dtype=torch.float32)
    g = g / g.sum()
    return g.to(dtype)


def l2pool(x: torch.Tensor) -> torch.Tensor:
    """sqrt(blur3x3(x^2, stride 2, pad 1) + 1e-12), depthwise (DISTS_pt.py:22-25, ADISTS.py:28-31)."""
    c = x.shape[1]
    filt = hanning_filter(x.dtype).to(x.device)[None, None].repeat(c, 1, 1, 1)
    out = F.conv2d(x * x, filt, stride=2, padding=1, groups=c)
    return (out + 1e-12).sqrt()


def trunk(x: torch.Tensor, conv_w: Sequence[torch.Tensor], conv_b: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """forward_once (DISTS_pt.py:91-103 == ADISTS.py:112-125): returns the six levels
    [raw x, relu1_2, relu2_2, relu3_3, relu4_3, relu5_3]."""
    mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    h = (x - mean) / std
    feats = [x]
    li = 0
    for stage, nconv in enumerate(STAGE_CONVS):
        if stage > 0:
            h = l2pool(h)
        for _ in range(nconv):
            h = F.relu(F.conv2d(h, conv_w[li].to(x.dtype), conv_b[li].to(x.dtype), stride=1, padding=1))
            li += 1
        feats.append(h)
    return feats


# ----------------------------------------------------------------------------------------------
# DISTS
# ----------------------------------------------------------------------------------------------
def dists_level_stats(fx: torch.Tensor, fy: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-(b,c) structure/texture similarity of one level (DISTS_pt.py:130-141): two-pass population
    variance, one-pass covariance, c1 = c2 = 1e-6.  Returns (S1, S2), each [B, C]."""
    c1 = c2 = 1e-6
    mx = fx.mean(dim=(2, 3), keepdim=True)
    my = fy.mean(dim=(2, 3), keepdim=True)
    s1 = (2 * mx * my + c1) / (mx * mx + my * my + c1)
    vx = ((fx - mx) ** 2).mean(dim=(2, 3), keepdim=True)
    vy = ((fy - my) ** 2).mean(dim=(2, 3), keepdim=True)
    cov = (fx * fy).mean(dim=(2, 3), keepdim=True) - mx * my
    s2 = (2 * cov + c2) / (vx + vy + c2)
    return s1.flatten(1), s2.flatten(1)


def dists_from_feats(feats_x: Sequence[torch.Tensor], feats_y: Sequence[torch.Tensor],
                     alpha: torch.Tensor, beta: torch.Tensor, batch_average: bool = False) -> torch.Tensor:
    """forward_from_feats / the statistics half of forward (DISTS_pt.py:113-148, 181-208)."""
    a = alpha.reshape(-1)
    b = beta.reshape(-1)
    w_sum = a.sum() + b.sum()
    a = a / w_sum
    b = b / w_sum
    total = 0
    off = 0
    for k, c in enumerate(CHNS):
        s1, s2 = dists_level_stats(feats_x[k], feats_y[k])
        total = total + (s1 * a[off:off + c]).sum(1) + (s2 * b[off:off + c]).sum(1)
        off += c
    score = 1 - total
    return score.mean() if batch_average else score


def dists_score(x, y, conv_w, conv_b, alpha, beta, batch_average: bool = False) -> torch.Tensor:
    """DISTS.forward (DISTS_pt.py:105-148)."""
    return dists_from_feats(trunk(x, conv_w, conv_b), trunk(y, conv_w, conv_b), alpha, beta, batch_average)


def project_weights(alpha: torch.Tensor, beta: torch.Tensor, lower_bound_rgb: float = 0.02):
    """DISTS.project_weights (DISTS_pt.py:82-89): floor the first three channels, renormalise jointly."""
    lb = torch.zeros_like(alpha)
    lb.reshape(-1)[:3] = lower_bound_rgb
    a = torch.maximum(alpha, lb)
    b = torch.maximum(beta, lb)
    s = a.sum() + b.sum()
    return a / s, b / s


# ----------------------------------------------------------------------------------------------
# A-DISTS
# ----------------------------------------------------------------------------------------------
def gaussian_window(window_size: int = 21, dtype=torch.float32) -> torch.Tensor:
    """2-D window = outer product of the normalised 1-D gaussian with sigma = window/3
    (ADISTS.py:64, 102-110)."""
    sigma = window_size / 3
    g = torch.tensor([math.exp(-(i - window_size // 2) ** 2 / float(2 * sigma ** 2)) for i in range(window_size)],
                     dtype=torch.float32)
    g = g / g.sum()
    return (g[:, None] @ g[None, :]).to(dtype)


def _win(t: torch.Tensor, win2d: torch.Tensor) -> torch.Tensor:
    c = t.shape[1]
    return F.conv2d(t, win2d[None, None].expand(c, 1, -1, -1).contiguous(), stride=1, padding=0, groups=c)


def _minmax_hw(t: torch.Tensor) -> torch.Tensor:
    lo = t.flatten(2).min(dim=-1, keepdim=True)[0].unsqueeze(-1)
    hi = t.flatten(2).max(dim=-1, keepdim=True)[0].unsqueeze(-1)
    return (t - lo) / (hi - lo + 1e-12)


def adists_structure_prob(feats_x: Sequence[torch.Tensor], window_size: int = 21) -> List[torch.Tensor]:
    """compute_prob (ADISTS.py:71-100): coarse-to-fine product of dispersion-index probability maps of
    the FIRST image's un-normalised features.  A level whose map is smaller than the window takes the
    global-statistics branch (the reference's bare ``except``)."""
    c0 = 1e-12
    x0 = feats_x[0]
    win = gaussian_window(window_size, x0.dtype).to(x0.device)
    prod = torch.ones_like(x0[:, :1])
    out: List[Optional[torch.Tensor]] = [None] * len(feats_x)
    for k in range(len(feats_x) - 1, -1, -1):
        f = feats_x[k]
        if f.shape[2] >= window_size and f.shape[3] >= window_size:
            m = _win(f, win)
            v = _win(f * f, win) - m * m
            gamma = (v / (m + c0)).mean(dim=1, keepdim=True)
            z = (gamma - gamma.mean(dim=(2, 3), keepdim=True)) / (gamma.std(dim=(2, 3), keepdim=True) + c0)
            ps = _minmax_hw(1 / (1 + torch.exp(-z)))
            prod = ps * F.interpolate(prod, size=m.shape[2:], mode="bilinear", align_corners=True)
            prod = _minmax_hw(prod)
        else:
            m = f.mean(dim=(2, 3), keepdim=True)
            v = ((f - m) ** 2).mean(dim=(2, 3), keepdim=True)
            gamma = (v / (m + c0)).mean(dim=1, keepdim=True)
            ps = 1 / (1 + torch.exp(-gamma))
            prod = ps * F.interpolate(prod, size=(1, 1), mode="bilinear", align_corners=True)
        out[k] = prod
    return out  # type: ignore[return-value]


def adists_channel_weights(feats_x: Sequence[torch.Tensor]) -> torch.Tensor:
    """entropy() per level + the joint normalise / clamp(mean +- 0.5 std) / renormalise
    (ADISTS.py:127-135, 152-160).  Returns [B, 1475, 1]."""
    c0 = 1e-12
    parts = []
    for f in feats_x:
        b, c = f.shape[:2]
        p = F.normalize(F.relu(f), dim=(2, 3)).reshape(b, c, -1)
        p = p / (p.sum(dim=2, keepdim=True) + c0)
        e = (-p * torch.log2(p + c0)).sum(dim=2, keepdim=True)
        e = e / (e.sum(dim=1, keepdim=True) + c0)
        parts.append(e * c)
    w = torch.cat(parts, dim=1)
    w = w / w.sum(dim=(1, 2), keepdim=True)
    mu = w.mean(dim=(1, 2), keepdim=True)
    sd = ((w - mu) ** 2).mean(dim=(1, 2), keepdim=True).sqrt()
    w = w.clamp(min=mu - 0.5 * sd, max=mu + 0.5 * sd)
    return w / w.sum(dim=(1, 2), keepdim=True)


def adists_from_feats(feats_x, feats_y, window_size: int = 21, return_parts: bool = False):
    """ADISTS.forward after the trunk (ADISTS.py:147-197), per-sample scores [B] (as_loss=False)."""
    ps = adists_structure_prob(feats_x, window_size)
    wts = torch.split(adists_channel_weights(feats_x), list(CHNS), dim=1)
    win = gaussian_window(window_size, feats_x[0].dtype).to(feats_x[0].device)
    d = 0
    per_level = []
    for k in range(len(CHNS) - 1, -1, -1):
        fx = F.normalize(feats_x[k], dim=(2, 3))
        fy = F.normalize(feats_y[k], dim=(2, 3))
        if fx.shape[2] >= window_size and fx.shape[3] >= window_size:
            mx, my = _win(fx, win), _win(fy, win)
            vx = _win(fx * fx, win) - mx * mx
            vy = _win(fy * fy, win) - my * my
            cov = _win(fx * fy, win) - mx * my
        else:
            mx = fx.mean(dim=(2, 3), keepdim=True)
            my = fy.mean(dim=(2, 3), keepdim=True)
            vx = ((fx - mx) ** 2).mean(dim=(2, 3), keepdim=True)
            vy = ((fy - my) ** 2).mean(dim=(2, 3), keepdim=True)
            cov = (fx * fy).mean(dim=(2, 3), keepdim=True) - mx * my
        t = (2 * mx * my + 1e-6) / (mx * mx + my * my + 1e-6)
        s = (2 * cov + 1e-6) / (vx + vy + 1e-6)
        p = ps[k].expand(mx.shape[0], mx.shape[1], -1, -1)
        dmap = (((1 - p) * t + p * s) * wts[k].unsqueeze(3)).sum(1, keepdim=True)
        lvl = dmap.mean(dim=(2, 3)).sum(1)
        per_level.append(lvl)
        d = d + lvl
    if return_parts:
        return 1 - d, ps, torch.cat(wts, dim=1), per_level[::-1]
    return 1 - d


def adists_score(x, y, conv_w, conv_b, window_size: int = 21) -> torch.Tensor:
    assert x.shape == y.shape                                  # ADISTS.py:138
    return adists_from_feats(trunk(x, conv_w, conv_b), trunk(y, conv_w, conv_b), window_size)


# ----------------------------------------------------------------------------------------------
# video-level pooling and correlations
# ----------------------------------------------------------------------------------------------
def pool_video(scores: np.ndarray) -> dict:
    """Temporal pooling of one video's per-frame scores (prep.py:191-198, logger.py:157-160)."""
    s = np.asarray(scores, dtype=np.float64)
    return {"mean": float(s.mean()), "std": float(s.std()), "min": float(s.min()), "max": float(s.max())}


def correlations(pred: np.ndarray, mos: np.ndarray) -> dict:
    """PLCC / SRCC / KTCC over videos (logger.py:93-102 -> scipy.stats)."""
    import scipy.stats as st
    return {"plcc": float(st.pearsonr(pred, mos)[0]), "srcc": float(st.spearmanr(pred, mos)[0]),
            "ktcc": float(st.kendalltau(pred, mos)[0])}
