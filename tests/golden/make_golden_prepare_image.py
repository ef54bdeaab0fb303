"""Golden vectors of the reference's prepare_image helpers (row a1 of SURVEY.md section 8): DISTS_pt.py:210-217,
DISTS_pt_original.py:140-144 and ADISTS.py:200-204, run UNMODIFIED from /root/reference on seeded PIL images.

    python tests/golden/make_golden_prepare_image.py  ->  tests/golden/prepare_image.npz   (build container only)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402,F401  (vgg16 shim + sys.path for the reference)
from PIL import Image  # noqa: E402

from nerf_qa.DISTS_pytorch.DISTS_pt import prepare_image as ref_dists_prepare  # noqa: E402

ref_adists_prepare = make_golden.ref_adists_mod.prepare_image


def image(h, w, seed):
    r = np.random.RandomState(seed)
    base = r.rand(h // 8 + 2, w // 8 + 2, 3)        # blocky content + a diagonal ramp: compresses well, resizes non-trivially
    img = np.kron(base, np.ones((8, 8, 1)))[:h, :w] * 200 + ((np.arange(h)[:, None, None] + np.arange(w)[None, :, None]) % 50)
    return img.astype(np.uint8)


def main():
    out = {}
    for name, (h, w, seed) in {"landscape": (270, 330, 1), "portrait": (340, 264, 2), "small": (120, 200, 3)}.items():
        u8 = image(h, w, seed)
        pil = Image.fromarray(u8, "RGB")
        out[f"{name}_input"] = u8
        # outputs are k / 255 exactly (ToTensor of a uint8 image): stored as uint8
        q = lambda t: np.round(t.numpy() * 255).astype(np.uint8)
        for key, t in (("dists_default", ref_dists_prepare(pil)), ("dists_keep_aspect", ref_dists_prepare(pil, keep_aspect_ratio=True)),
                       ("dists_noresize", ref_dists_prepare(pil, resize=False)), ("adists_default", ref_adists_prepare(pil))):
            assert np.array_equal(q(t).astype(np.float32) / 255, t.numpy())
            if key in ("dists_default", "dists_keep_aspect"):
                out[f"{name}_{key}"] = q(t)
            else:       # equal to the input / the keep-aspect output: a digest is enough
                out[f"{name}_{key}_shape"] = np.array(t.shape)
                out[f"{name}_{key}_sha256"] = np.frombuffer(hashlib.sha256(q(t).tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "prepare_image.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
