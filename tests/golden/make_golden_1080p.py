"""Golden vectors at the BENCHMARKED shape: one 1920x1080 frame pair through the UNMODIFIED reference
DISTS and ADISTS modules (same shims as make_golden.py; fp32 CPU, the reference's own arithmetic).

    python tests/golden/make_golden_1080p.py        # ~2 minutes of CPU in the build container

Written: tests/golden/pair_1080p.npz with the two per-frame scores and the A-DISTS intermediates the GPU
tests compare level by level (rows a6-a8 of SURVEY.md section 8): the 1475 entropy channel weights after
clamping (ADISTS.py:152-160), the coarse ps maps in full and 8x-decimated samples of the fine ones
(ADISTS.py:71-100), and the per-level contributions D_k (ADISTS.py:165-191).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (patches torchvision.models.vgg16 and imports the reference modules)
from nerf_qa_b200 import synth  # noqa: E402

SEED = 31   # the pair tests/test_gpu_parity.py::test_dists_1080p_against_oracle uses


def adists_levels(m, x, y):
    """ADISTS.forward (ADISTS.py:147-191) re-run statement by statement on the reference module's own methods so that
    the per-level terms can be recorded; the final score is asserted equal to the module's forward."""
    feats0, feats1 = m.forward_once(x), m.forward_once(y)
    ps_list = m.compute_prob(feats0)
    c1 = c2 = 1e-6
    pad = torch.nn.ReflectionPad2d(0)
    weight = torch.concat([m.entropy(f) for f in feats0], dim=1)
    weight = weight / weight.sum(dim=(1, 2), keepdim=True)
    mean = weight.mean(dim=(1, 2), keepdim=True)
    stdv = ((weight - mean) ** 2).mean(dim=(1, 2), keepdim=True).sqrt()
    weight = weight.clamp(min=mean - 0.5 * stdv, max=mean + 0.5 * stdv)
    weight = weight / weight.sum(dim=(1, 2), keepdim=True)
    w_levels = torch.split(weight, m.chns, dim=1)
    d_levels = [None] * 6
    for k in range(5, -1, -1):
        ps = ps_list[k].expand(x.shape[0], m.chns[k], -1, -1)
        fx, fy = F.normalize(feats0[k], dim=(2, 3)), F.normalize(feats1[k], dim=(2, 3))
        win = m.windows[k]
        xm = F.conv2d(pad(fx), win, stride=1, padding=0, groups=fx.shape[1])
        ym = F.conv2d(pad(fy), win, stride=1, padding=0, groups=fx.shape[1])
        xv = F.conv2d(pad(fx * fx), win, stride=1, padding=0, groups=fx.shape[1]) - xm * xm
        yv = F.conv2d(pad(fy * fy), win, stride=1, padding=0, groups=fx.shape[1]) - ym * ym
        cov = F.conv2d(pad(fx * fy), win, stride=1, padding=0, groups=fx.shape[1]) - xm * ym
        t = (2 * xm * ym + c1) / (xm ** 2 + ym ** 2 + c1)
        s = (2 * cov + c2) / (xv + yv + c2)
        dmap = (((1 - ps) * t + ps * s) * w_levels[k].unsqueeze(3)).sum(1, keepdim=True)
        d_levels[k] = dmap.mean([2, 3]).sum(1)
    return ps_list, weight.squeeze(-1), torch.stack(d_levels, dim=1)


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 1)
    x, y = synth.make_frame_pairs(1, 1080, 1920, seed=SEED)
    d32, a32 = G.build_ref_dists(torch.float32), G.build_ref_adists(torch.float32)
    with torch.no_grad():
        sd = d32(x, y)
        sa = a32(x, y, as_loss=False)
        ps, wt, dk = adists_levels(a32, x, y)
    assert abs(float(1 - dk.sum()) - float(sa)) < 2e-6, (float(1 - dk.sum()), float(sa))
    out = dict(
        shape=np.array([1, 1080, 1920, SEED]), dists_f32=sd.numpy(), adists_f32=sa.numpy(),
        adists_weights=wt.numpy(), adists_level_terms=dk.numpy(),
        ps_shapes=np.array([list(p.shape) for p in ps]),
        ps_summary=np.stack([np.array([float(p.double().mean()), float(p.double().std()), float(p.min()), float(p.max())])
                             for p in ps]),
        ps_level5=ps[5].numpy().copy(), ps_level4=ps[4].numpy().copy(), ps_level3=ps[3].numpy().copy(),
        ps_level2_dec8=ps[2][..., ::8, ::8].numpy().copy(), ps_level1_dec8=ps[1][..., ::8, ::8].numpy().copy(),
        ps_level0_dec8=ps[0][..., ::8, ::8].numpy().copy(),
        input_checksum=np.array([float(x.double().sum()), float(y.double().sum())]),
    )
    np.savez_compressed(os.path.join(HERE, "pair_1080p.npz"), **out)
    print("DISTS", sd.numpy(), "A-DISTS", sa.numpy(), "levels", dk.numpy())


if __name__ == "__main__":
    main()
