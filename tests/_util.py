"""Shared helpers for the GPU parity tests."""
import warnings

import numpy as np
import torch

from nerf_qa_b200 import synth

VGG_SEED = 1234


def load_synth_vgg(module):
    ws, bs = synth.make_vgg_weights(VGG_SEED)
    sd = module.state_dict()
    for i, idx in enumerate(synth.VGG_FEATURE_IDX):
        st = synth.CONV_STAGE[i]
        sd[f"stage{st}.{idx}.weight"] = ws[i]
        sd[f"stage{st}.{idx}.bias"] = bs[i]
    module.load_state_dict(sd)
    return module


def make_dists(precision, device="cuda:0"):
    import nerf_qa_b200
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = nerf_qa_b200.DISTS(precision=precision)
    return load_synth_vgg(m).to(device).eval()


def make_adists(precision, device="cuda:0"):
    import nerf_qa_b200
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = nerf_qa_b200.ADISTS(precision=precision)
    return load_synth_vgg(m).to(device).eval()


def golden_inputs(g):
    b, h, w, seed = (int(v) for v in g["shape"])
    return synth.make_frame_pairs(b, h, w, seed=seed)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12)))


def report(name, got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    print(f"[parity] {name}: max|abs| {np.max(np.abs(got - want)):.3e}  max rel {rel_err(got, want):.3e}")
