"""Row f2: the resize variants of the reference's loaders (test2_prep.py:109-114, 221-227, 324-327, 424-439; prep.py:91-96)
on the GPU.  The oracle is the reference's own call: torch's F.interpolate on the CPU for the plain bilinear variants, and
torchvision's resize of a PIL image (+ ToTensor) for the antialiased ones (equal up to PIL's uint8 rounding)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from nerf_qa_b200 import frames


def _ref_pixel_count_size(oh_in, ow_in):
    # test2_prep.py:426-436, verbatim arithmetic
    OH, OW = oh_in, ow_in
    if OW >= OH:
        ratio = float(OW) / float(OH)
        H = math.sqrt(256 * 256 / ratio)
        W = int(ratio * H)
    else:
        ratio = float(OH) / float(OW)
        W = math.sqrt(256 * 256 / ratio)
        H = int(ratio * W)
    return H, W


@pytest.mark.parametrize("hw", [(1080, 1920), (800, 800), (540, 960), (378, 504), (200, 300), (1920, 1080)])
def test_target_sizes_follow_the_reference_loaders(hw):
    h, w = hw
    H, W = _ref_pixel_count_size(h, w)
    # the reference hands the float H (or W) to F.interpolate; this torch wants ints, so the float is truncated here
    x = torch.zeros(1, 3, h, w)
    want = F.interpolate(x, size=(int(H), int(W)), mode="bilinear", align_corners=False).shape[2:]
    assert frames.target_size("pixel_count", h, w)[:2] == tuple(want)
    assert abs(want[0] * want[1] - 256 * 256) <= 256 * 3          # "pixel count": H * W ~ 65536
    from PIL import Image
    from torchvision import transforms
    img = Image.new("RGB", (w, h))
    if min(img.size) > 256:
        a = transforms.functional.resize(img, 256)
        s = transforms.functional.resize(img, (256, 256))
    else:
        a = s = img
    assert frames.target_size("aspect", h, w)[:2] == (a.size[1], a.size[0])
    assert frames.target_size("square", h, w)[:2] == (s.size[1], s.size[0])
    assert frames.target_size("full_size", h, w)[:2] == (h, w)
    assert frames.target_size("interp256", h, w) == (256, 256, False)


def _image(h, w, seed):
    g = torch.Generator().manual_seed(seed)
    base = F.interpolate(torch.rand(1, 3, h // 7 + 2, w // 7 + 2, generator=g), size=(h, w), mode="bicubic").clamp(0, 1)
    noisy = (base + 0.1 * torch.rand(1, 3, h, w, generator=g)).clamp(0, 1)
    return (noisy * 255).round().to(torch.uint8)            # what a decoded PNG holds


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(1080, 1920), (378, 504), (801, 400), (256, 300)])
def test_resize_variants_against_the_reference_calls(hw):
    from PIL import Image
    from torchvision import transforms
    h, w = hw
    u8 = _image(h, w, seed=h + w)
    x = u8.float() / 255.0                                   # transforms.ToTensor()
    xs = x.cuda()
    # plain bilinear: the literal reference op (test2_prep.py:438, prep.py:94-96)
    for mode in ("pixel_count", "interp256"):
        oh, ow, aa = frames.target_size(mode, h, w)
        assert not aa
        want = F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False)
        got = frames.prepare_frames(xs, mode).cpu()
        err = float((got - want).abs().max())
        print(f"[parity] resize {mode} {hw}->{(oh, ow)}: max|abs| {err:.3e}")
        assert got.shape == want.shape and err < 2e-6
    # antialiased: torch's aa kernel on the CPU (same algorithm, fp32) and the PIL path the reference runs
    pil = Image.fromarray(u8[0].permute(1, 2, 0).numpy(), "RGB")
    for mode in ("aspect", "square"):
        oh, ow, aa = frames.target_size(mode, h, w)
        got = frames.prepare_frames(xs, mode).cpu()
        if not aa:
            assert torch.equal(got, x)
            continue
        want_t = F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False, antialias=True)
        ref = transforms.functional.resize(pil, 256 if mode == "aspect" else (256, 256))
        want_p = transforms.ToTensor()(ref).unsqueeze(0)
        e_t, e_p = float((got - want_t).abs().max()), float((got - want_p).abs().max())
        print(f"[parity] resize {mode} {hw}->{(oh, ow)}: vs torch aa {e_t:.3e}, vs PIL {e_p:.3e} (uint8 step {1 / 255:.3e})")
        assert got.shape == want_p.shape
        assert e_t < 3e-5                # same taps; fp32 summation order differs from ATen's
        assert e_p < 1.01 / 255          # PIL rounds to uint8 after each of its two passes
    assert frames.prepare_frames(xs, "full_size") is xs


@pytest.mark.gpu
def test_resized_frames_feed_the_scorer():
    """Upload once, resize on the device, score: the scores equal scoring the CPU-resized frames (same fp32 values)."""
    from tests._util import make_dists
    m = make_dists("fp32")
    u8a, u8b = _image(540, 960, 1), _image(540, 960, 2)
    xa, xb = u8a.float() / 255, u8b.float() / 255
    with torch.no_grad():
        got = m(frames.prepare_frames(xa.cuda(), "pixel_count"), frames.prepare_frames(xb.cuda(), "pixel_count")).cpu()
        size = frames.target_size("pixel_count", 540, 960)[:2]
        want = m(F.interpolate(xa, size=size, mode="bilinear", align_corners=False).cuda(),
                 F.interpolate(xb, size=size, mode="bilinear", align_corners=False).cuda()).cpu()
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-6)
