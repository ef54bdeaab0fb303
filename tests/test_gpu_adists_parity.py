"""A-DISTS parity at the benchmarked shape and level by level (rows a6-a8 of SURVEY.md section 8), the host-buffer
entry points, and the stream-ordering contract of a context.  Everything goes through the C-ABI.

Tolerances (BASELINE.json north_star): fp32 path 1e-4 relative, bf16 path 2e-3 absolute on per-frame scores."""
import numpy as np
import pytest
import torch

from nerf_qa_b200 import _lib, synth
from oracle import reference_oracle as O
from tests._util import golden_inputs, make_adists, make_dists, rel_err, report

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


@pytest.fixture(scope="module")
def amodels():
    return {p: make_adists(p) for p in ("fp32", "bf16")}


@pytest.fixture(scope="module")
def dmodels():
    return {p: make_dists(p) for p in ("fp32", "bf16")}


def _ctx_parts(model, x, y):
    ctx = model._context(x.device)
    return ctx.adists_parts(x, y, model.window_size, model._prec())


def test_adists_1080p_against_reference_golden(amodels, dmodels, golden):
    """BASELINE configs[2] shape.  tests/golden/pair_1080p.npz holds what the UNMODIFIED reference modules return
    for this pair (tests/golden/make_golden_1080p.py) -- scores, the 1475 channel weights, the six per-level terms
    D_k and the ps maps; every level is windowed at this size (1060x1900 ... 48x100 maps)."""
    g = golden("pair_1080p")
    x, y = golden_inputs(g)
    xs, ys = x.cuda(), y.cuda()
    for prec in ("fp32", "bf16"):
        score, wts, terms, ps = _ctx_parts(amodels[prec], xs, ys)
        got = score.cpu().numpy()
        report(f"A-DISTS 1080p {prec}", got, g["adists_f32"])
        report(f"A-DISTS 1080p {prec} level terms", terms.cpu().numpy(), g["adists_level_terms"])
        report(f"A-DISTS 1080p {prec} channel weights", wts.cpu().numpy(), g["adists_weights"])
        plain = amodels[prec](xs, ys, as_loss=False).cpu().numpy()
        assert np.array_equal(plain, got)                      # the parts entry point is the same computation
        full = {5: g["ps_level5"], 4: g["ps_level4"], 3: g["ps_level3"]}
        dec = {2: g["ps_level2_dec8"], 1: g["ps_level1_dec8"], 0: g["ps_level0_dec8"]}
        ps_err = 0.0
        for k in range(6):
            assert tuple(ps[k].shape) == tuple(g["ps_shapes"][k])
            want = full[k] if k in full else dec[k]
            have = ps[k].cpu().numpy() if k in full else ps[k][..., ::8, ::8].cpu().numpy()
            ps_err = max(ps_err, float(np.max(np.abs(have - want))))
        print(f"[parity] A-DISTS 1080p {prec} ps maps: max|abs| {ps_err:.3e}")
        if prec == "fp32":
            # this pair scores 0.0044 = 1 - 0.9956: 1e-4 relative is 4e-7 absolute, at the rounding floor of the fp32 sum of
            # the six level terms (each ~0.04-0.34, ulp 3e-8), hence the absolute alternative
            assert rel_err(got, g["adists_f32"]) < 1e-4 or np.max(np.abs(got - g["adists_f32"])) < 1e-6
            assert np.max(np.abs(terms.cpu().numpy() - g["adists_level_terms"])) < 2e-5
            assert rel_err(wts.cpu().numpy(), g["adists_weights"]) < 1e-4
            assert ps_err < 1e-3          # maps in [0,1] after two min-max normalisations of a sigmoid
        else:
            assert np.max(np.abs(got - g["adists_f32"])) < 2e-3
            assert np.max(np.abs(terms.cpu().numpy() - g["adists_level_terms"])) < 2e-3
            assert ps_err < 5e-2
    # the DISTS half of the headline metric against the same reference run
    for prec, tol_rel, tol_abs in (("fp32", 1e-4, None), ("bf16", None, 2e-3)):
        got = dmodels[prec](xs, ys).cpu().numpy()
        report(f"DISTS 1080p {prec} (reference golden)", got, g["dists_f32"])
        if tol_rel:
            assert rel_err(got, g["dists_f32"]) < tol_rel
        else:
            assert np.max(np.abs(got - g["dists_f32"])) < tol_abs


@pytest.mark.parametrize("name", ["adists_96x128", "adists_176", "adists_256", "adists_352x368"])
def test_adists_intermediates_against_oracle(amodels, golden, vgg_weights, name):
    """compute_prob maps (a6), entropy weights (a7) and per-level terms (a8) on the GPU against the oracle (which
    tests/test_oracle_golden.py pins to the reference's own ps maps and entropies), windowed and fallback levels."""
    g = golden(name)
    x, y = golden_inputs(g)
    ws, bs = vgg_weights
    fx, fy = O.trunk(x, ws, bs), O.trunk(y, ws, bs)
    want_s, want_ps, want_w, want_terms = O.adists_from_feats(fx, fy, return_parts=True)
    score, wts, terms, ps = _ctx_parts(amodels["fp32"], x.cuda(), y.cuda())
    assert rel_err(score.cpu().numpy(), want_s.numpy()) < 1e-4
    assert rel_err(wts.cpu().numpy(), want_w.squeeze(-1).numpy()) < 1e-4
    want_t = torch.stack(want_terms, dim=1).numpy()
    report(f"A-DISTS {name} level terms", terms.cpu().numpy(), want_t)
    assert np.max(np.abs(terms.cpu().numpy() - want_t)) < 2e-5
    for k in range(6):
        assert tuple(ps[k].shape) == tuple(want_ps[k].shape), f"level {k}"
        err = float((ps[k].cpu() - want_ps[k]).abs().max())
        print(f"[parity] A-DISTS {name} ps level {k} {tuple(ps[k].shape)}: max|abs| {err:.3e}")
        assert err < 1e-3, f"level {k}"
    # golden summary of the reference's own maps
    for k in range(6):
        p = ps[k].double()
        mean, lo, hi = float(p.mean()), float(p.min()), float(p.max())
        assert abs(mean - g["ps_summary"][k][0]) < 1e-3 and abs(lo - g["ps_summary"][k][2]) < 1e-3 and abs(hi - g["ps_summary"][k][3]) < 1e-3


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_adists_host_path_equals_device_path(amodels, dmodels, prec):
    """nqa_adists_forward_host / nqa_dists_adists_forward_host (frames streamed from pinned host memory in micro-batches,
    first chunk shortened) give the device path's scores bit for bit: every reduction is batch-independent."""
    x, y = synth.make_frame_pairs(5, 176, 208, seed=71)
    xs, ys = x.cuda(), y.cuda()
    a, d = amodels[prec], dmodels[prec]
    want = a(xs, ys, as_loss=False).cpu()
    got = a.score_host(x.pin_memory(), y.pin_memory())
    assert not got.is_cuda and torch.equal(got, want)
    dd, aa = d.score_with_adists(xs, ys)
    hd, ha = d.score_with_adists_host(x.pin_memory(), y.pin_memory())
    assert torch.equal(hd, dd.cpu()) and torch.equal(ha, aa.cpu())
    with pytest.raises(_lib.NerfQAError):
        a.score_host(xs, ys)                        # host entry point, device tensors


def test_forced_small_micro_batches_on_the_host_path(monkeypatch):
    """With NQA_MICRO_IMAGES forcing one-pair micro-batches the host path walks its double-buffered staging ring several
    times (slot reuse waits on the consumed events) and still reproduces the device path."""
    monkeypatch.setenv("NQA_MICRO_IMAGES", "1")
    x, y = synth.make_frame_pairs(7, 540, 960, seed=73)
    a = make_adists("bf16")
    d = make_dists("bf16")
    assert a._context(torch.device("cuda:0")).micro_batch(7, 540, 960, a._prec()) <= 2
    want = a(x.cuda(), y.cuda(), as_loss=False).cpu()
    assert torch.equal(a.score_host(x.pin_memory(), y.pin_memory()), want)
    assert torch.equal(d.score_host(x.pin_memory(), y.pin_memory()), d(x.cuda(), y.cuda()).cpu())


def test_calls_on_different_streams_are_ordered(dmodels):
    """One context, several streams: a device call left pending on a side stream, then the host path (its own
    internal stream), then a call on the default stream -- all share the workspace arena and the channel weights, and
    the context orders them with its event.  alpha/beta produced by torch ops on the side stream must be the ones
    the scoring call on that stream reads (stream-ordered nqa_set_alpha_beta_async)."""
    m = make_dists("bf16")
    x, y = synth.make_frame_pairs(6, 270, 480, seed=79)
    xs, ys = x.cuda(), y.cuda()
    want = m(xs, ys).clone()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    for _ in range(3):
        with torch.cuda.stream(side):
            s_side = m(xs, ys)                                       # pending on the side stream
        s_host = m.score_host(x.pin_memory(), y.pin_memory())        # internal compute stream, same arena
        s_main = m(xs, ys)                                           # default stream
        torch.cuda.synchronize()
        assert torch.equal(s_side, want) and torch.equal(s_host, want.cpu()) and torch.equal(s_main, want)
    # weights written by a kernel on the side stream right before the call
    a0 = m.alpha.data.clone()
    big = torch.rand(64, 1475, 1475, device="cuda")                  # keeps the side stream busy ahead of the edit
    with torch.cuda.stream(side):
        for _ in range(4):
            big = torch.bmm(big, big).clamp_(0, 1)
        m.alpha.data = (a0 * 3.0 + big[0, 0, 0] * 0.0)               # depends on the long-running work
        s_scaled = m(xs, ys)
    torch.cuda.synchronize()
    ref = make_dists("bf16")
    ref.alpha.data = a0 * 3.0
    assert torch.equal(s_scaled, ref(xs, ys))
    m.alpha.data = a0


def test_data_edits_of_alpha_beta_are_seen(dmodels):
    """In-place writes through .data bump no version counter; alpha/beta are uploaded on every call, and
    invalidate_weights() covers the VGG tensors."""
    m = make_dists("bf16")
    x, y = synth.make_frame_pairs(2, 64, 64, seed=83)
    xs, ys = x.cuda(), y.cuda()
    s0 = m(xs, ys).clone()
    m.alpha.data.mul_(0.5)
    s1 = m(xs, ys).clone()
    assert not torch.equal(s0, s1)
    m.alpha.data.mul_(2.0)
    assert torch.equal(m(xs, ys), s0)
    conv = m.stage1._modules["0"]
    w0 = conv.weight.data.clone()
    conv.weight.data.mul_(1.5)
    m.invalidate_weights()
    assert not torch.equal(m(xs, ys), s0)
    conv.weight.data.copy_(w0)
    m.invalidate_weights()
    assert torch.equal(m(xs, ys), s0)


def test_paths_without_a_backward_refuse_gradients(amodels, dmodels):
    x, y = synth.make_frame_pairs(1, 64, 64, seed=89)
    xs = x.cuda().requires_grad_(True)
    with torch.enable_grad():
        with pytest.raises(NotImplementedError):
            amodels["fp32"](xs, y.cuda())                 # the reference's default as_loss=True is a loss
        with pytest.raises(NotImplementedError):
            dmodels["fp32"].forward_once(xs)
        with pytest.raises(NotImplementedError):
            dmodels["fp32"].stage2._modules["4"](torch.rand(1, 64, 8, 8, device="cuda", requires_grad=True))
    assert amodels["fp32"](xs.detach(), y.cuda()).dim() == 0


@pytest.mark.parametrize("shape", [(2, 352, 368), (3, 176, 208), (1, 540, 960)])
def test_strip_plane_path_against_the_planar_path(monkeypatch, shape):
    """bf16 trunk: by default the last convolution of a stage writes its output as strip planes from the epilogue and one
    kernel pools + sums them (NQA_ADISTS_STRIPS=2).  Against it: strip planes transposed from the NHWC level (1: validates
    the strip readers alone) and plain planar planes (0: the round-1 kernels).  Same trunk arithmetic; the pooling kernels
    round the pooled activations differently in rare last bits (MUFU sqrt of separable sums vs sqrtf of nine fmas), hence a
    tolerance far inside the bf16 budget instead of equality.  Widths 208 -> 104 -> 52: levels that cannot be strips
    (w % 8 != 0) fall back per level.  DISTS from the shared-trunk call must agree too."""
    b, h, w = shape
    x, y = synth.make_frame_pairs(b, h, w, seed=83)
    xs, ys = x.cuda(), y.cuda()
    res = {}
    for mode in ("2", "1", "0"):
        monkeypatch.setenv("NQA_ADISTS_STRIPS", mode)
        a, d = make_adists("bf16"), make_dists("bf16")
        res[mode] = (a(xs, ys, as_loss=False).cpu().numpy(), *[t.cpu().numpy() for t in d.score_with_adists(xs, ys)])
    for mode in ("1", "0"):
        for got, want, name in zip(res["2"], res[mode], ("A-DISTS", "DISTS (shared trunk)", "A-DISTS (shared trunk)")):
            report(f"strips=2 vs strips={mode} {name} {shape}", got, want)
            assert np.max(np.abs(got - want)) < 1e-4
    # the strip readers alone (same pooled activations as the planar path: only the layout of the planes differs)
    assert np.array_equal(res["1"][0], res["0"][0])


@pytest.mark.parametrize("shape", [(1, 2160, 3840), (2, 1088, 1928)])
def test_large_and_unaligned_frames_on_the_strip_path(monkeypatch, shape):
    """4K frames (32-bit plane strides, 17280-element tensor-map rows) and a width whose deeper levels stop being multiples
    of 8 (1928 -> 964 -> 482 ...: strip planes for level 1 only, planar below): the strip-plane path against the planar one."""
    b, h, w = shape
    x, y = synth.make_frame_pairs(b, h, w, seed=5)
    xs, ys = x.cuda(), y.cuda()
    out = {}
    for mode in ("2", "0"):
        monkeypatch.setenv("NQA_ADISTS_STRIPS", mode)
        a, d = make_adists("bf16"), make_dists("bf16")
        out[mode] = (a(xs, ys, as_loss=False).cpu().numpy(), *[t.cpu().numpy() for t in d.score_with_adists(xs, ys)])
        assert np.all(np.isfinite(out[mode][0])) and np.all(out[mode][0] > 0)
        del a, d
        torch.cuda.empty_cache()
    for got, want in zip(out["2"], out["0"]):
        assert np.max(np.abs(got - want)) < 1e-4


def test_adists_full_size_properties(amodels, dmodels):
    """Size-independent properties at the benchmark's frame size (1080p), across a micro-batch boundary (9 pairs = 8 + 1):
    a frame's A-DISTS score does not depend on the batch it travels in (bit for bit, alone or through the shared-trunk
    call's A-DISTS half), identical frames score 0 (T = S = 1 and the entropy weights sum to one), and the host path
    returns the device path's scores."""
    x, y = synth.make_frame_pairs(9, 1080, 1920, seed=101)
    xs, ys = x.cuda(), y.cuda()
    a, d = amodels["bf16"], dmodels["bf16"]
    s = a(xs, ys, as_loss=False)
    assert torch.equal(a(xs[:1], ys[:1], as_loss=False), s[:1])
    assert torch.equal(a(xs[8:], ys[8:], as_loss=False), s[8:])
    assert torch.equal(a(xs[3:6], ys[3:6], as_loss=False), s[3:6])
    both = d.score_with_adists(xs, ys)[1]
    assert torch.equal(d.score_with_adists(xs[8:], ys[8:])[1], both[8:])
    assert torch.allclose(both, s, rtol=1e-4, atol=1e-5)
    same = a(xs[:2], xs[:2], as_loss=False)
    assert float(same.abs().max()) < 1e-5
    assert torch.equal(a.score_host(x.pin_memory(), y.pin_memory()), s.cpu())
