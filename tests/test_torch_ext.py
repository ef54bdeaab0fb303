"""The thin PyTorch C++ extension over the C-ABI (nerf_qa_b200/csrc_torch/nqa_torch_ops.cpp, torch.ops.nerfqa_b200.*), the
binding BASELINE.json's north_star names.  CPU: it loads, registers every operator with a fake kernel, and has no CPU
kernel.  GPU: the nn.Modules reach the library through it and give the ctypes binding's results bit for bit."""
import pytest
import torch

from nerf_qa_b200 import _lib, synth, torch_ops
from tests._util import make_adists, make_dists


def test_extension_loads_and_registers_every_operator():
    assert torch_ops.available(), "nerf_qa_b200/_nqa_torch_ops.so missing: `make` builds it"
    ops = torch_ops.load()
    for name in torch_ops.OPS:
        assert hasattr(ops, name), name
    # nothing behind the operators on the CPU: the dispatcher refuses, loudly
    x = torch.rand(1, 3, 32, 32)
    with pytest.raises(NotImplementedError):
        ops.dists_forward(1, x, x, _lib.PREC_BF16, True)
    with pytest.raises(NotImplementedError):
        ops.l2pool(1, x)


def test_fake_kernels_give_the_output_shapes():
    from torch._subclasses.fake_tensor import FakeTensorMode
    ops = torch_ops.load()
    with FakeTensorMode():
        x = torch.empty(3, 3, 45, 77, device="cuda")
        s, s12 = ops.dists_forward(1, x, x, _lib.PREC_BF16, True)
        assert s.shape == (3,) and s12.shape == (3, 2, _lib.NUM_CHANNELS)
        assert ops.dists_forward(1, x, x, _lib.PREC_BF16, False)[1].shape[0] == 0
        assert ops.adists_forward(1, x, x, 21, _lib.PREC_BF16).shape == (3,)
        d, a = ops.dists_adists_forward(1, x, x, 21, _lib.PREC_BF16)
        assert d.shape == a.shape == (3,)
        assert [tuple(t.shape) for t in ops.dists_features(1, x, _lib.PREC_FP32)] == list(_lib.level_shapes(3, 45, 77)[1:])
        assert ops.l2pool(1, torch.empty(2, 64, 45, 77, device="cuda")).shape == (2, 64, 23, 39)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_modules_through_the_extension_equal_the_ctypes_binding(monkeypatch, prec):
    x, y = synth.make_frame_pairs(3, 96, 136, seed=77)
    x, y = x.cuda(), y.cuda()
    f = torch.rand(2, 64, 33, 47, device="cuda")
    with torch.no_grad():
        d, a = make_dists(prec), make_adists(prec)
        assert d._context(x.device).ops is not None, "the extension is built: the modules must use it"
        got = (d(x, y), a(x, y, as_loss=False), *d.score_with_adists(x, y), *d.forward_once(x)[1:],
               d.stage2._modules["4"](f))
        monkeypatch.setenv("NQA_BINDING", "ctypes")
        dc, ac = make_dists(prec), make_adists(prec)
        assert dc._context(x.device).ops is None
        want = (dc(x, y), ac(x, y, as_loss=False), *dc.score_with_adists(x, y), *dc.forward_once(x)[1:],
                dc.stage2._modules["4"](f))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert torch.equal(g, w)


@pytest.mark.gpu
def test_operators_called_directly_and_their_errors():
    ops = torch_ops.load()
    x, y = synth.make_frame_pairs(2, 64, 64, seed=5)
    x, y = x.cuda(), y.cuda()
    with torch.no_grad():
        d = make_dists("bf16")
        handle = int(d._context(x.device).handle.value)
        d(x, y)                                             # uploads the weights
        s, s12 = ops.dists_forward(handle, x, y, _lib.PREC_BF16, True)
        assert torch.equal(s, d(x, y)) and s12.shape == (2, 2, _lib.NUM_CHANNELS)
        with pytest.raises(RuntimeError, match="shape mismatch"):
            ops.dists_forward(handle, x, y[:, :, :32], _lib.PREC_BF16, True)
        with pytest.raises(RuntimeError, match="nerfqa_b200 error"):
            ops.adists_forward(handle, x, y, 22, _lib.PREC_BF16)        # even window: rejected by the library
        # a context of its own, created and destroyed through the operators
        h2 = ops.ctx_create(0)
        with pytest.raises(RuntimeError, match="nerfqa_b200 error"):      # no weights yet
            ops.dists_forward(h2, x, y, _lib.PREC_BF16, False)
        ops.ctx_destroy(h2)
