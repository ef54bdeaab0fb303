"""Row a1: prepare_image (DISTS_pt.py:210-217, ADISTS.py:200-204) against outputs of the reference's own helpers on the
same PIL images (tests/golden/make_golden_prepare_image.py).  Host code, no GPU."""
import hashlib

import numpy as np
import pytest
from PIL import Image

import nerf_qa_b200
from nerf_qa_b200 import adists, dists


@pytest.mark.parametrize("name", ["landscape", "portrait", "small"])
def test_prepare_image_equals_the_reference(golden, name):
    g = golden("prepare_image")
    pil = Image.fromarray(g[f"{name}_input"], "RGB")
    cases = {
        "dists_default": dists.prepare_image(pil),
        "dists_keep_aspect": dists.prepare_image(pil, keep_aspect_ratio=True),
        "dists_noresize": dists.prepare_image(pil, resize=False),
        "adists_default": adists.prepare_image(pil),
    }
    for key, got in cases.items():
        assert got.dtype.is_floating_point and float(got.min()) >= 0.0 and float(got.max()) <= 1.0
        q = np.round(got.numpy() * 255).astype(np.uint8)          # ToTensor output is k / 255 exactly
        assert np.array_equal(q.astype(np.float32) / 255, got.numpy()), key
        if f"{name}_{key}" in g:                                   # stored as uint8
            assert q.shape == g[f"{name}_{key}"].shape and np.array_equal(q, g[f"{name}_{key}"]), key
        else:                                                      # stored as a digest
            assert list(q.shape) == g[f"{name}_{key}_shape"].tolist(), key
            assert hashlib.sha256(q.tobytes()).digest() == g[f"{name}_{key}_sha256"].tobytes(), key
    assert nerf_qa_b200.prepare_image is dists.prepare_image
    # the contract the scorers rely on: [1,3,H,W], short side <= 256 after the default resize
    assert cases["dists_default"].shape[0] == 1 and cases["dists_default"].shape[1] == 3
    assert min(cases["dists_keep_aspect"].shape[2:]) <= 256
