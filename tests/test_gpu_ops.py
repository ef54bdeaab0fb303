"""Op-level parity on the GPU: each hand-written kernel against plain torch fp32 / the CPU oracle.
Everything goes through the C-ABI (nerf_qa_b200._lib.Context)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from nerf_qa_b200 import _lib
from oracle import reference_oracle as O
from tests._util import report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _lib.Context(torch.device("cuda:0"))


def _conv_case(b, cin, cout, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, cin, h, w, generator=g)          # post-ReLU-like, non-negative
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    return x, wt, bias


CONV_SHAPES = [
    (1, 64, 64, 16, 8),      # exactly one 128-pixel tile, BLOCK_N = 64
    (2, 64, 128, 37, 53),    # ragged edges, BLOCK_N = 128
    (1, 128, 256, 24, 40),   # BLOCK_N = 256, two k chunks
    (2, 128, 128, 40, 21),   # two 64-channel tiles with resident weights (halo kernel, conv2_2 shape class)
    (2, 256, 512, 17, 30),   # two channel tiles per pixel tile
    (1, 512, 512, 9, 15),    # deep layer, 72 k-blocks, image smaller than a tile
    (3, 64, 64, 135, 240),   # many tiles per CTA (persistent loop, TMEM double buffering)
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3x3_fp32_kernel(ctx, shape):
    x, w, b = _conv_case(*shape, seed=1)
    want = F.relu(F.conv2d(x, w, b, padding=1))
    got = ctx.conv3x3_relu(x.cuda(), w.cuda(), b.cuda(), _lib.PREC_FP32).cpu()
    report(f"conv fp32 {shape}", got, want)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3x3_tcgen05_bf16(ctx, shape):
    x, w, b = _conv_case(*shape, seed=2)
    # the kernel rounds operands to bf16 and accumulates in fp32: compare with exactly that
    xq, wq = x.bfloat16().float(), w.bfloat16().float()
    want = F.relu(F.conv2d(xq, wq, b, padding=1)).bfloat16().float()
    got = ctx.conv3x3_relu(x.cuda(), w.cuda(), b.cuda(), _lib.PREC_BF16).cpu()
    report(f"conv tcgen05 bf16 {shape}", got, want)
    # one bf16 ulp (2^-8 relative) of slack for accumulation-order differences at rounding boundaries
    torch.testing.assert_close(got, want, rtol=8e-3, atol=1e-3)


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3x3_tcgen05_bf16x3(ctx, shape):
    x, w, b = _conv_case(*shape, seed=3)
    want = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1)).float()
    got = ctx.conv3x3_relu(x.cuda(), w.cuda(), b.cuda(), _lib.PREC_BF16X3).cpu()
    report(f"conv tcgen05 bf16x3 {shape}", got, want)
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("shape", [(2, 3, 9, 13), (1, 64, 135, 240), (2, 5, 1, 1), (1, 7, 2, 31)])
def test_l2pool_nchw(ctx, shape):
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(5)) * 3
    want = O.l2pool(x)
    got = ctx.l2pool(x.cuda()).cpu()
    report(f"l2pool {shape}", got, want)
    torch.testing.assert_close(got, want, rtol=2e-6, atol=1e-7)


def _feats(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    out_x, out_y = [], []
    for (bb, c, hh, ww) in _lib.level_shapes(b, h, w):
        fx = torch.rand(bb, c, hh, ww, generator=g) * (0.5 if c == 3 else 2.0)
        fy = (fx + 0.3 * torch.randn(fx.shape, generator=g)).clamp_min(0)
        out_x.append(fx)
        out_y.append(fy)
    return out_x, out_y


@pytest.mark.parametrize("shape", [(2, 64, 64), (3, 45, 77), (1, 135, 240)])
def test_dists_from_feats_matches_oracle(ctx, shape, alpha_beta):
    fx, fy = _feats(*shape, seed=7)
    a, b = alpha_beta
    ctx.set_alpha_beta(a, b)
    want = O.dists_from_feats(fx, fy, a, b)
    score, s1s2 = ctx.dists_from_feats([f.cuda() for f in fx], [f.cuda() for f in fy])
    s1 = torch.cat([O.dists_level_stats(fx[k], fy[k])[0] for k in range(6)], 1)
    s2 = torch.cat([O.dists_level_stats(fx[k], fy[k])[1] for k in range(6)], 1)
    report(f"from_feats S1 {shape}", s1s2[:, 0].cpu(), s1)
    report(f"from_feats S2 {shape}", s1s2[:, 1].cpu(), s2)
    report(f"from_feats score {shape}", score.cpu(), want)
    torch.testing.assert_close(s1s2[:, 0].cpu(), s1, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(s1s2[:, 1].cpu(), s2, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(score.cpu(), want, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("shape", [(2, 96, 128), (1, 176, 176), (1, 352, 368)])
def test_adists_from_feats_matches_oracle(ctx, shape):
    fx, fy = _feats(*shape, seed=9)
    want, ps, wts, per_level = O.adists_from_feats(fx, fy, return_parts=True)
    got = ctx.adists_from_feats([f.cuda() for f in fx], [f.cuda() for f in fy], 21).cpu()
    report(f"adists from_feats {shape}", got, want)
    # white-noise "features" are the worst case for the windowed variance cancellation (the oracle's own
    # fp32-vs-fp64 distance is of this order); the golden end-to-end cases hold 1e-4
    torch.testing.assert_close(got, want, rtol=3e-4, atol=1e-6)
