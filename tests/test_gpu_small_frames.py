"""The reference's own call pattern: 256 x 256 frames at batch 4 (nerf_qa/settings.py:1, prep.py:163).  Such calls are
launch-bound, so the library replays a CUDA graph captured at the second call of a shape; the replay must give the plain
path's scores bit for bit, follow weight updates, and keep the caller's stream order."""
import json
import os
import time

import numpy as np
import pytest
import torch

from nerf_qa_b200 import synth
from oracle import reference_oracle as O
from tests._util import make_adists, make_dists, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _plain(make, monkeypatch):
    monkeypatch.setenv("NQA_GRAPHS", "0")
    m = make("bf16")
    m._context(torch.device("cuda:0"))
    monkeypatch.delenv("NQA_GRAPHS")
    return m


def test_graph_replay_equals_the_plain_path(monkeypatch):
    x, y = synth.make_frame_pairs(4, 256, 256, seed=91)
    xs, ys = x.cuda(), y.cuda()
    d, a = make_dists("bf16"), make_adists("bf16")
    dp, ap = _plain(make_dists, monkeypatch), _plain(make_adists, monkeypatch)
    want_d, want_a = dp(xs, ys).clone(), ap(xs, ys, as_loss=False).clone()
    ctx = d._context(xs.device)
    for it in range(5):
        x2, y2 = xs.clone(), ys.clone()                      # fresh addresses every call: the graph runs on staging buffers
        assert torch.equal(d(x2, y2), want_d), it
        assert torch.equal(a(x2, y2, as_loss=False), want_a), it
        dd, aa = d.score_with_adists(x2, y2)
        assert torch.allclose(dd, want_d, rtol=1e-5, atol=1e-7), it    # moments of the strip-plane levels: another fixed order
        assert torch.equal(dd, dp.score_with_adists(x2, y2)[0]), it    # replay == plain path of the same entry point
    assert ctx.graph_replays() >= 4 and dp._context(xs.device).graph_replays() == 0
    # other inputs through the same graph
    x3, y3 = synth.make_frame_pairs(4, 256, 256, seed=92)
    assert torch.equal(d(x3.cuda(), y3.cuda()), dp(x3.cuda(), y3.cuda()))
    # weights are read at replay time: alpha / beta edits and a new VGG state are seen
    d.alpha.data.mul_(0.5)
    dp.alpha.data.mul_(0.5)
    assert torch.equal(d(xs, ys), dp(xs, ys))
    conv = d.stage3._modules["12"]
    conv.weight.data.mul_(1.1)
    d.invalidate_weights()
    dp.stage3._modules["12"].weight.data.mul_(1.1)
    dp.invalidate_weights()
    assert torch.equal(d(xs, ys), dp(xs, ys))
    # a larger call in between regrows the workspace and drops the graphs; the small shape is captured again
    big = synth.make_frame_pairs(2, 540, 960, seed=93)
    d(big[0].cuda(), big[1].cuda())
    for _ in range(3):
        assert torch.equal(d(xs, ys), dp(xs, ys))


def test_graph_replay_keeps_the_stream_order():
    d = make_dists("bf16")
    x, y = synth.make_frame_pairs(4, 256, 256, seed=94)
    xs, ys = x.cuda(), y.cuda()
    want = d(xs, ys).clone()
    d(xs, ys)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    big = torch.rand(32, 1024, 1024, device="cuda")
    for _ in range(3):
        with torch.cuda.stream(side):
            for _ in range(3):
                big = torch.bmm(big, big).clamp_(0, 1)
            xin = xs + big[0, 0, 0] * 0.0                   # produced late on the side stream
            s = d(xin, ys)
            t = s * 2.0                                      # consumer on the same stream right after
        torch.cuda.synchronize()
        assert torch.equal(s, want) and torch.equal(t, want * 2.0)


def test_256_batch_4_matches_the_oracle(vgg_weights, alpha_beta):
    """BASELINE configs[0] shape family through the graph path, against the oracle (fp32: 1e-4 relative)."""
    x, y = synth.make_frame_pairs(4, 256, 256, seed=95)
    ws, bs = vgg_weights
    a, b = alpha_beta
    want_d = O.dists_score(x, y, ws, bs, a, b).numpy()
    want_a = O.adists_score(x, y, ws, bs).numpy()
    d32, a32 = make_dists("fp32"), make_adists("fp32")
    for _ in range(3):
        got_d, got_a = d32(x.cuda(), y.cuda()).cpu().numpy(), a32(x.cuda(), y.cuda(), as_loss=False).cpu().numpy()
    assert rel_err(got_d, want_d) < 1e-4 and rel_err(got_a, want_a) < 1e-4
    d16, a16 = make_dists("bf16"), make_adists("bf16")
    for _ in range(3):
        g_d, g_a = d16(x.cuda(), y.cuda()).cpu().numpy(), a16(x.cuda(), y.cuda(), as_loss=False).cpu().numpy()
    assert np.max(np.abs(g_d - want_d)) < 2e-3 and np.max(np.abs(g_a - want_a)) < 2e-3


def test_eager_gpu_reference_timing_and_tf32_floor(vgg_weights, alpha_beta):
    """Context numbers, not a target: the reference's own GPU path is eager PyTorch over cuDNN.
    The oracle port on CUDA tensors is timed with TF32 convolutions on (the torch default the reference runs with) and
    off, and the TF32-vs-fp32 score distance is recorded -- the parity floor of the reference's GPU mode.  Written to
    gpurun_out/eager_gpu_baseline.json when that directory exists."""
    ws, bs = vgg_weights
    a, b = alpha_beta
    dev = torch.device("cuda:0")
    wsd, bsd, ad, bd = [w.to(dev) for w in ws], [t.to(dev) for t in bs], a.to(dev), b.to(dev)
    out = {}
    for name, (bsz, h, w) in {"256x256_b4": (4, 256, 256), "1080p_b2": (2, 1080, 1920)}.items():
        x, y = synth.make_frame_pairs(bsz, h, w, seed=96)
        xd, yd = x.to(dev), y.to(dev)
        res = {}
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for kind, fn in (("dists", lambda: O.dists_score(xd, yd, wsd, bsd, ad, bd)),
                             ("adists", lambda: O.adists_score(xd, yd, wsd, bsd))):
                fn()
                torch.cuda.synchronize()
                n = 5 if h <= 256 else 2
                t0 = time.perf_counter()
                for _ in range(n):
                    s = fn()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / n
                res[f"{kind}_{'tf32' if tf32 else 'fp32'}"] = {"pairs_per_s": bsz / dt, "ms_per_call": dt * 1e3,
                                                               "scores": s.cpu().tolist()}
        torch.backends.cudnn.allow_tf32 = True
        for kind in ("dists", "adists"):
            t32, f32 = np.array(res[f"{kind}_tf32"]["scores"]), np.array(res[f"{kind}_fp32"]["scores"])
            res[f"{kind}_tf32_vs_fp32_max_abs"] = float(np.max(np.abs(t32 - f32)))
            res[f"{kind}_tf32_vs_fp32_max_rel"] = float(np.max(np.abs(t32 - f32) / np.abs(f32)))
            print(f"[eager-gpu] {name} {kind}: TF32 {res[f'{kind}_tf32']['pairs_per_s']:.1f} pairs/s, fp32 "
                  f"{res[f'{kind}_fp32']['pairs_per_s']:.1f} pairs/s, TF32-vs-fp32 |abs| {res[f'{kind}_tf32_vs_fp32_max_abs']:.2e}")
            assert res[f"{kind}_tf32_vs_fp32_max_abs"] < 2e-3      # the reference's own GPU mode sits inside the bf16 tolerance
        out[name] = res
    gdir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(gdir):
        with open(os.path.join(gdir, "eager_gpu_baseline.json"), "w") as f:
            json.dump(out, f, indent=1)
