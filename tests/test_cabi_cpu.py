"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol the header declares,
fails loudly without a GPU, and the Python modules keep the reference's surface."""
import ctypes
import os
import re
import warnings

import pytest
import torch

import nerf_qa_b200
from nerf_qa_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "nerfqa_b200.h")).read()
    return sorted(set(re.findall(r"NQA_API\s+[\w\s\*]+?\b(nqa_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = nerf_qa_b200.load_library()
    declared = _header_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nerfqa_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.nqa_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = nerf_qa_b200.load_library()
    h = ctypes.c_void_p()
    rc = lib.nqa_ctx_create(0, ctypes.byref(h))
    assert rc == -2 and b"no CUDA device" in lib.nqa_last_error()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = nerf_qa_b200.DISTS()
        a = nerf_qa_b200.ADISTS()
    x = torch.rand(1, 3, 32, 32)
    with pytest.raises(_lib.NerfQAError):
        m(x, x)
    with pytest.raises(_lib.NerfQAError):
        a(x, x)
    with pytest.raises(_lib.NerfQAError):
        m.forward_once(x)
    with pytest.raises(_lib.NerfQAError):
        nerf_qa_b200.L2pooling(channels=3)(x)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nerf_qa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read(), f"{f} mentions the oracle"


def test_state_dict_layout_matches_reference():
    """34 keys with the reference's names, shapes of torchvision's VGG16."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = nerf_qa_b200.DISTS()
    sd = m.state_dict()
    assert len(sd) == 34
    assert sd["alpha"].shape == (1, 1475, 1, 1) and sd["beta"].shape == (1, 1475, 1, 1)
    assert abs(float(sd["alpha"].sum() + sd["beta"].sum()) - 1.0) < 1e-5      # shipped weights are normalised
    assert sd["mean"].flatten().tolist() == pytest.approx([0.485, 0.456, 0.406])
    for key, shape in {"stage1.0.weight": (64, 3, 3, 3), "stage1.2.bias": (64,), "stage2.4.filter": (64, 1, 3, 3),
                       "stage3.9.filter": (128, 1, 3, 3), "stage3.14.weight": (256, 256, 3, 3),
                       "stage4.16.filter": (256, 1, 3, 3), "stage5.23.filter": (512, 1, 3, 3),
                       "stage5.28.weight": (512, 512, 3, 3)}.items():
        assert tuple(sd[key].shape) == shape, key
    assert all(not p.requires_grad for n, p in m.named_parameters() if n.startswith("stage"))
    assert m.alpha.requires_grad and m.beta.requires_grad                     # created after the freeze loop
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = nerf_qa_b200.ADISTS()
    asd = a.state_dict()
    assert "windows.0" in asd and asd["windows.5"].shape == (512, 1, 21, 21) and "stage2.4.filter" in asd
    assert abs(float(asd["windows.1"][0].sum()) - 1.0) < 1e-5


def test_project_weights_matches_oracle():
    from oracle import reference_oracle as O
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = nerf_qa_b200.DISTS()
    m.alpha.data = m.alpha.data - 0.001
    a0, b0 = m.alpha.data.clone(), m.beta.data.clone()
    m.project_weights()
    pa, pb = O.project_weights(a0, b0)
    torch.testing.assert_close(m.alpha.data, pa)
    torch.testing.assert_close(m.beta.data, pb)


def test_alpha_beta_backward_formula():
    """The autograd glue around the kernel's (score, S1, S2): closed-form gradients w.r.t. the effective channel
    weights, chained by autograd through the weight normalisation, equal autograd through the plain expression
    (DISTS_pt.py:127-144).  The library context is replaced by a stand-in that evaluates the same formula."""
    from nerf_qa_b200.dists import _WeightedScore
    g = torch.Generator().manual_seed(0)
    s1s2 = torch.rand(3, 2, 1475, generator=g, dtype=torch.float64)

    class FakeCtx:
        def set_channel_weights(self, a, b):
            self.a, self.b = a.detach().flatten(), b.detach().flatten()

        def dists_forward(self, x, y, prec, want_s1s2=True):
            return 1 - ((s1s2[:, 0] * self.a).sum(1) + (s1s2[:, 1] * self.b).sum(1)), s1s2

    alpha = torch.rand(1, 1475, 1, 1, generator=g, dtype=torch.float64, requires_grad=True)
    beta = torch.rand(1, 1475, 1, 1, generator=g, dtype=torch.float64, requires_grad=True)
    w = alpha.sum() + beta.sum()
    plain = 1 - ((s1s2[:, 0] * alpha.flatten()).sum(1) + (s1s2[:, 1] * beta.flatten()).sum(1)) / w
    coef = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    ga, gb = torch.autograd.grad((plain * coef).sum(), [alpha, beta])
    x = torch.zeros(3, 3, 4, 4, dtype=torch.float64)
    out = _WeightedScore.apply(x, x, alpha / w, beta / w, FakeCtx(), 0, False)
    torch.testing.assert_close(out, plain.detach())
    ga2, gb2 = torch.autograd.grad((out * coef).sum(), [alpha, beta])
    torch.testing.assert_close(ga2, ga)
    torch.testing.assert_close(gb2, gb)


def test_autograd_glue_routes_frame_gradients_through_the_activation_tape():
    """_WeightedScore keeps the activations (dists_forward_saved) only when a frame needs a gradient through the trunk,
    consumes the tape in the first backward, recomputes in a second backward through a retained graph, and falls
    back to recomputation when the library declines the tape.  The library context is a stand-in that records calls."""
    from nerf_qa_b200.dists import _WeightedScore
    from nerf_qa_b200 import _lib

    class Tape:
        def __init__(self):
            self.id = 7

    class FakeCtx:
        def __init__(self, give_tape):
            self.calls, self.give_tape = [], give_tape

        def set_channel_weights(self, a, b):
            pass

        def _out(self, x):
            return torch.zeros(x.shape[0]), torch.zeros(x.shape[0], 2, 1475)

        def dists_forward(self, x, y, prec, want_s1s2=True):
            self.calls.append("forward")
            return self._out(x)

        def dists_forward_saved(self, x, y, prec):
            self.calls.append("forward_saved")
            return (*self._out(x), Tape() if self.give_tape else None)

        def dists_backward_saved(self, tape, x, y, g):
            self.calls.append("backward_saved")
            tape.id = 0
            return torch.ones_like(x), torch.ones_like(y)

        def dists_backward(self, x, y, prec, g, through_trunk):
            self.calls.append(f"backward_recompute(through_trunk={through_trunk})")
            return torch.ones_like(x), torch.ones_like(y)

    a = torch.rand(1, 1475, 1, 1)
    b = torch.rand(1, 1475, 1, 1)

    def run(ctx, through_trunk, frames_need_grad, backwards=1):
        x = torch.zeros(2, 3, 4, 4, requires_grad=frames_need_grad)
        y = torch.zeros(2, 3, 4, 4)
        w = a.clone().requires_grad_(True)
        out = _WeightedScore.apply(x, y, w, b, ctx, _lib.PREC_BF16, through_trunk)
        for i in range(backwards):
            out.sum().backward(retain_graph=i + 1 < backwards)
        return ctx.calls

    assert run(FakeCtx(True), True, True) == ["forward_saved", "backward_saved"]
    assert run(FakeCtx(True), True, True, backwards=2) == ["forward_saved", "backward_saved",
                                                           "backward_recompute(through_trunk=True)"]
    assert run(FakeCtx(False), True, True) == ["forward_saved", "backward_recompute(through_trunk=True)"]
    # the reference's default (trunk under no_grad): level-0 gradient only, nothing to keep
    assert run(FakeCtx(True), False, True) == ["forward", "backward_recompute(through_trunk=False)"]
    # only the channel weights need gradients: closed form from S1 / S2, no backward call into the library
    assert run(FakeCtx(True), True, False) == ["forward"]


def _build_c_example(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "score_frames")
    libdir = os.path.join(root, "nerf_qa_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "score_frames.c"), "-L", libdir, "-lnerfqa_b200",
                    f"-Wl,-rpath,{libdir}", "-lm", "-o", exe], check=True)
    return exe


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """include/nerfqa_b200.h compiles as strict C99 and examples/score_frames.c links against the shared library; without
    a GPU the program must stop at context creation with the no-fallback error (exit code 2), never compute on the CPU."""
    import subprocess
    import torch
    exe = _build_c_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2, (r.stdout, r.stderr)
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_consumer_scores_frames_through_the_c_abi(tmp_path):
    """The plain-C consumer scores host frames with nqa_dists_forward_host and checks DISTS(x,x) = 0 and symmetry."""
    import subprocess
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe, "3", "135", "240"], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert r.stdout.strip().endswith("OK")
