"""FR model wrapper (nerf_qa/model_stats.py NeRFQAModel) against goldens generated from the reference module
(tests/golden/make_golden_frmodel.py): head initialisation and entropy regulariser on CPU, forward + one training
step on the GPU."""
import warnings

import numpy as np
import pytest
import torch

from nerf_qa_b200.fr_model import NeRFQAModel
from nerf_qa_b200 import synth

CASES = {
    "frmodel_linear_relu": dict(regression_type="linear", subjective_score_type="MOS", dists_weight_norm="relu"),
    "frmodel_sqrt_off": dict(regression_type="sqrt", subjective_score_type="MOS", dists_weight_norm="off"),
    "frmodel_logistic_softmax": dict(regression_type="logistic", subjective_score_type="DMOS", dists_weight_norm="softmax"),
}


def build(g, cfg, precision="fp32"):
    target = g["train_dmos"] if cfg["subjective_score_type"] == "DMOS" else g["train_mos"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return NeRFQAModel(dists=g["train_dists"], scores=target, precision=precision, weight_lower_bound=1e-4,
                           alpha_beta_ratio=2.0, detach_beta="False", **cfg)


@pytest.mark.parametrize("name", sorted(CASES))
def test_head_init_and_entropy_match_reference(golden, name):
    g = golden(name)
    m = build(g, CASES[name])
    for k, v in zip(g["head_names"], g["head_values"]):
        assert abs(float(getattr(m, str(k))) - float(v)) <= 1e-4 * max(1.0, abs(float(v))), k
    assert abs(float(m.entropy_loss()) - float(g["entropy_loss"])) < 1e-4 * abs(float(g["entropy_loss"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_forward_and_training_step(golden, name):
    from tests._util import load_synth_vgg
    g = golden(name)
    m = build(g, CASES[name])
    load_synth_vgg(m.dists_model)
    m = m.cuda()
    x, y = synth.make_frame_pairs(3, 64, 64, seed=29)
    scores, dists_scores = m(x.cuda(), y.cuda())
    np.testing.assert_allclose(dists_scores.detach().cpu().numpy(), g["dists_scores"], rtol=1e-4)
    np.testing.assert_allclose(scores.detach().cpu().numpy(), g["scores"], rtol=2e-4)
    # one optimisation step of the FR training loop (run_final.py:189-219): L1 to a target + entropy regulariser
    params = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)
    target = torch.tensor([3.0, 2.0, 4.0], device="cuda")
    loss = (scores - target).abs().mean() + 0.1 * m.entropy_loss()
    opt.zero_grad()
    loss.backward()
    assert m.dists_model.alpha.grad is not None and float(m.dists_model.alpha.grad.abs().max()) > 0
    assert all(p.grad is not None for p in params if p is not m.dists_model.beta or CASES[name]["dists_weight_norm"] != "x")
    before = scores.detach().clone()
    opt.step()
    after, _ = m(x.cuda(), y.cuda())
    assert not torch.equal(after.detach(), before)
