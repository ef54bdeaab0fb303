"""Golden vectors for the FR model wrapper (nerf_qa/model_stats.py NeRFQAModel) from the UNMODIFIED reference module
(build container only; shims as in make_golden_flavours.py)."""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import pandas as pd
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from nerf_qa_b200 import synth  # noqa: E402

_orig_vgg16 = torchvision.models.vgg16
torchvision.models.vgg16 = lambda *a, **k: _orig_vgg16(weights=None)
_orig_load = torch.load
_BLOB = "/root/reference/nerf_qa/DISTS_pytorch/weights.pt"
torch.load = lambda f, *a, **k: _orig_load(_BLOB if str(f).endswith("weights.pt") else f, *a, **k)

import wandb  # noqa: E402

CASES = {
    "frmodel_linear_relu": dict(regression_type="linear", subjective_score_type="MOS", dists_weight_norm="relu"),
    "frmodel_sqrt_off": dict(regression_type="sqrt", subjective_score_type="MOS", dists_weight_norm="off"),
    "frmodel_logistic_softmax": dict(regression_type="logistic", subjective_score_type="DMOS", dists_weight_norm="softmax"),
}
COMMON = dict(weight_lower_bound=1e-4, alpha_beta_ratio=2.0, detach_beta="False")


def train_table():
    rng = np.random.default_rng(7)
    d = rng.uniform(0.02, 0.3, size=40)
    mos = 4.6 - 9.0 * d + rng.normal(0, 0.15, size=40)
    return pd.DataFrame({"DISTS": d, "MOS": mos, "DMOS": 100 - 20 * mos})


def main() -> None:
    import importlib
    df = train_table()
    x, y = synth.make_frame_pairs(3, 64, 64, seed=29)
    ws, bs = synth.make_vgg_weights(1234)
    for name, cfg in CASES.items():
        wandb.config = types.SimpleNamespace(**cfg, **COMMON)
        for modname in ("DISTS_pt_original", "DISTS_pt_softmax"):
            importlib.import_module(f"nerf_qa.DISTS_pytorch.{modname}").device = torch.device("cpu")
        ms = importlib.import_module("nerf_qa.model_stats")
        ms.device = torch.device("cpu")
        m = ms.NeRFQAModel(df)
        sd = m.dists_model.state_dict()
        for i, idx in enumerate(synth.VGG_FEATURE_IDX):
            sd[f"stage{synth.CONV_STAGE[i]}.{idx}.weight"] = ws[i]
            sd[f"stage{synth.CONV_STAGE[i]}.{idx}.bias"] = bs[i]
        m.dists_model.load_state_dict(sd)
        m.eval()
        scores, dists_scores = m(x, y)
        ent = m.entropy_loss()
        head = {k: float(getattr(m, k)) for k in ("dists_weight", "dists_bias", "b1", "b2", "b3", "b4") if hasattr(m, k)}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), train_dists=df["DISTS"].values, train_mos=df["MOS"].values,
                            train_dmos=df["DMOS"].values, scores=scores.detach().numpy(), dists_scores=dists_scores.detach().numpy(),
                            entropy_loss=np.array(float(ent)), head_names=np.array(list(head)), head_values=np.array(list(head.values())),
                            config=np.array([f"{k}={v}" for k, v in cfg.items()]))
        print(name, head, scores.detach().numpy(), float(ent))


if __name__ == "__main__":
    main()
