"""Golden vectors for ADISTS.forward(as_map=True) from the UNMODIFIED reference module (build container only;
same shims as make_golden.py)."""
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
import nerf_qa.ADISTS  # noqa: E402,F401

mod = sys.modules["nerf_qa.ADISTS.ADISTS"]
mod.device = torch.device("cpu")

CASES = {"adists_map_96x128": (96, 128, 4), "adists_map_176": (176, 176, 2)}   # H, W, seed (levels fall back / windowed mix)


def main() -> None:
    m = mod.ADISTS()
    ws, bs = synth.make_vgg_weights(1234)
    sd = m.state_dict()
    for i, idx in enumerate(synth.VGG_FEATURE_IDX):
        sd[f"stage{synth.CONV_STAGE[i]}.{idx}.weight"] = ws[i]
        sd[f"stage{synth.CONV_STAGE[i]}.{idx}.bias"] = bs[i]
    m.load_state_dict(sd)
    m.eval()
    for name, (h, w, seed) in CASES.items():
        x, y = synth.make_frame_pairs(2, h, w, seed=seed)
        with torch.no_grad():
            one = m(x[:1], y[:1], as_loss=False, as_map=True)          # [1,1,H,W]
            two = m(x, y, as_loss=False, as_map=True)                  # [2,2,H,W] (upstream broadcast quirk)
            score = m(x[:1], y[:1], as_loss=False)
        assert torch.allclose(two[0, 0], one[0, 0], atol=1e-6) and torch.equal(two[0, 0], two[0, 1])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array([2, h, w, seed]), map_b1=one.numpy(),
                            map_b2_shape=np.array(two.shape), map_b2_sample1=two[1, 0].numpy(), score_b1=score.numpy())
        print(name, tuple(one.shape), tuple(two.shape), float(one.mean()), float(score))


if __name__ == "__main__":
    main()
