"""Generate golden vectors by running the UNMODIFIED reference modules from /root/reference.

Run in the build container only (``python tests/golden/make_golden.py``); /root/reference does not
exist on the GPU box, so the outputs are committed as small .npz fixtures next to this script.

Shims, none of which touch reference source:
  * ``torchvision.models.vgg16`` is patched to build the architecture without the ImageNet download
    (offline container), then the seeded synthetic weights of ``nerf_qa_b200.synth.make_vgg_weights``
    are loaded through the reference's own state_dict keys (``stage{k}.{idx}.weight``);
  * ``DISTS(load_weights=False)`` + alpha/beta assigned from the reference's shipped
    ``nerf_qa/DISTS_pytorch/weights.pt`` (the file the reference expects under ``sys.prefix``).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from nerf_qa_b200 import synth  # noqa: E402

_orig_vgg16 = torchvision.models.vgg16
torchvision.models.vgg16 = lambda *a, **k: _orig_vgg16(weights=None)

from nerf_qa.DISTS_pytorch.DISTS_pt import DISTS as RefDISTS  # noqa: E402
import nerf_qa.ADISTS  # noqa: E402,F401  (the package re-exports the class over the submodule name)

ref_adists_mod = sys.modules["nerf_qa.ADISTS.ADISTS"]

ref_adists_mod.device = torch.device("cpu")
RefADISTS = ref_adists_mod.ADISTS

VGG_SEED = 1234


def load_synth_vgg(module: torch.nn.Module, dtype=torch.float32) -> None:
    ws, bs = synth.make_vgg_weights(VGG_SEED)
    sd = module.state_dict()
    for i, idx in enumerate(synth.VGG_FEATURE_IDX):
        st = synth.CONV_STAGE[i]
        sd[f"stage{st}.{idx}.weight"] = ws[i].to(dtype)
        sd[f"stage{st}.{idx}.bias"] = bs[i].to(dtype)
    module.load_state_dict(sd)


def build_ref_dists(dtype=torch.float32) -> torch.nn.Module:
    m = RefDISTS(load_weights=False)
    blob = torch.load("/root/reference/nerf_qa/DISTS_pytorch/weights.pt")
    m.alpha.data = blob["alpha"].clone()
    m.beta.data = blob["beta"].clone()
    m = m.to(dtype)
    load_synth_vgg(m, dtype)
    return m.eval()


def build_ref_adists(dtype=torch.float32) -> torch.nn.Module:
    m = RefADISTS().to(dtype)
    load_synth_vgg(m, dtype)
    return m.eval()


# name -> (B, H, W, seed)
DISTS_CASES = {
    "dists_256": (5, 256, 256, 0),        # BASELINE config 1 shape, one frame per distortion type
    "dists_odd_45x77": (3, 45, 77, 3),    # odd sizes through every L2pool (45->23->12->6->3)
    "dists_135x240": (2, 135, 240, 5),    # the 135 -> 68 pooling edge of the 1080p pyramid
    "dists_64": (5, 64, 64, 11),
}
ADISTS_CASES = {
    "adists_256": (5, 256, 256, 0),       # level 5 is 16x16 -> global-statistics fallback
    "adists_176": (2, 176, 176, 2),       # level 4 = 22 (windowed), level 5 = 11 (fallback)
    "adists_96x128": (2, 96, 128, 4),     # levels 3,4,5 fall back (24x32 level 3 is windowed: 24>=21)
    "adists_352x368": (1, 352, 368, 6),   # level 5 = 22x23 -> every level windowed
}


def summarize(feats):
    return np.stack([np.array([float(f.double().mean()), float(f.double().std()), float(f.double().abs().max())])
                     for f in feats])


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    d32, d64 = build_ref_dists(torch.float32), build_ref_dists(torch.float64)
    for name, (b, h, w, seed) in DISTS_CASES.items():
        x, y = synth.make_frame_pairs(b, h, w, seed=seed)
        with torch.no_grad():
            s32 = d32(x, y)
            s64 = d64(x.double(), y.double())
            fx, fy = d32.forward_once(x), d32.forward_once(y)
            sff = d32.forward_from_feats(fx, fy)
            savg = d32(x, y, batch_average=True)
        assert torch.equal(s32, sff)
        # per-(b,c) S1/S2 recomputed with the reference's exact expressions (DISTS_pt.py:130-141)
        s1s, s2s = [], []
        for k in range(6):
            xm = fx[k].mean([2, 3], keepdim=True); ym = fy[k].mean([2, 3], keepdim=True)
            s1s.append(((2 * xm * ym + 1e-6) / (xm ** 2 + ym ** 2 + 1e-6)).flatten(1))
            xv = ((fx[k] - xm) ** 2).mean([2, 3], keepdim=True); yv = ((fy[k] - ym) ** 2).mean([2, 3], keepdim=True)
            cov = (fx[k] * fy[k]).mean([2, 3], keepdim=True) - xm * ym
            s2s.append(((2 * cov + 1e-6) / (xv + yv + 1e-6)).flatten(1))
        out[name] = dict(
            shape=np.array([b, h, w, seed]), score_f32=s32.numpy(), score_f64=s64.numpy(),
            score_batch_avg=np.array(float(savg)), s1=torch.cat(s1s, 1).numpy(), s2=torch.cat(s2s, 1).numpy(),
            feat_summary_x=summarize(fx), feat_summary_y=summarize(fy),
            level_shapes=np.array([list(f.shape) for f in fx]),
            input_checksum=np.array([float(x.double().sum()), float(y.double().sum()), float(x[0, 1, h // 2, w // 3]),
                                     float(y[-1, 2, h // 3, w // 2])]),
            relu5_3_x0=fx[5][0, :8].numpy().copy(),
        )
        print(name, s32.numpy(), "fp64 delta", np.abs(s32.numpy() - s64.numpy()).max())

    a32, a64 = build_ref_adists(torch.float32), build_ref_adists(torch.float64)
    for name, (b, h, w, seed) in ADISTS_CASES.items():
        x, y = synth.make_frame_pairs(b, h, w, seed=seed)
        with torch.no_grad():
            s32 = a32(x, y, as_loss=False)
            s64 = a64(x.double(), y.double(), as_loss=False)
            sloss = a32(x, y)                      # default as_loss=True -> scalar 1 - mean(D)
            fx = a32.forward_once(x)
            ps = a32.compute_prob(fx)
            wt = torch.concat([a32.entropy(f) for f in fx], dim=1)
            syx = a32(y, x, as_loss=False)         # asymmetric: ps / weights come from the first argument
        out[name] = dict(
            shape=np.array([b, h, w, seed]), score_f32=s32.numpy(), score_f64=s64.numpy(),
            score_as_loss=np.array(float(sloss)), score_swapped=syx.numpy(),
            entropy_raw=wt.squeeze(-1).numpy(),
            ps_shapes=np.array([list(p.shape) for p in ps]),
            ps_summary=np.stack([np.array([float(p.double().mean()), float(p.double().std()), float(p.min()), float(p.max())])
                                 for p in ps]),
            ps_coarse=ps[3].numpy().copy(), ps_level5=ps[5].numpy().copy(),
            input_checksum=np.array([float(x.double().sum()), float(y.double().sum())]),
        )
        print(name, s32.numpy(), "fp64 delta", np.abs(s32.numpy() - s64.numpy()).max())

    for name, d in out.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print("wrote", len(out), "fixtures to", HERE)


if __name__ == "__main__":
    main()
