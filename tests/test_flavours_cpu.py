"""Host-side logic of the reference's weight-normalisation flavours (DISTS_pt_original.py / DISTS_pt_softmax.py)
against golden vectors generated from the reference modules (tests/golden/make_golden_flavours.py).  No GPU needed:
only the alpha/beta handling is exercised here; scores and gradients are checked in test_gpu_parity.py."""
import warnings

import numpy as np
import pytest
import torch

import nerf_qa_b200

CASES = {
    "flavour_original_relu": (nerf_qa_b200.DISTSOriginal, dict(weight_lower_bound=1e-4, alpha_beta_ratio=2.0,
                                                               dists_weight_norm="relu", detach_beta="False")),
    "flavour_original_relu_wsumdetach_detachbeta": (nerf_qa_b200.DISTSOriginal,
                                                    dict(weight_lower_bound=1e-3, alpha_beta_ratio=3.0,
                                                         dists_weight_norm="relu+w_sum_detach", detach_beta="True")),
    "flavour_softmax": (nerf_qa_b200.DISTSSoftmax, dict(detach_beta="False")),
}


def build(name, **extra):
    cls, cfg = CASES[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return cls(**cfg, **extra)


@pytest.mark.parametrize("name", sorted(CASES))
def test_loaded_and_projected_weights_match_reference(golden, name):
    g = golden(name)
    m = build(name)
    np.testing.assert_allclose(m.alpha.detach().flatten().numpy(), g["loaded_alpha"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(m.beta.detach().flatten().numpy(), g["loaded_beta"], rtol=1e-6, atol=0)
    a, b = m._effective_weights()
    assert abs(float(a.sum() + b.sum()) - 1.0) < 1e-5           # every flavour hands normalised weights to the kernels
    m.project_weights()
    np.testing.assert_allclose(m.alpha.detach().flatten().numpy(), g["projected_alpha"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(m.beta.detach().flatten().numpy(), g["projected_beta"], rtol=1e-5, atol=1e-9)


def test_detach_options_cut_the_graph():
    m = build("flavour_original_relu_wsumdetach_detachbeta")
    a, b = m._effective_weights()
    assert a.requires_grad and not b.requires_grad
    (a.sum() * 3.0).backward()
    # with w_sum detached, d(sum a / w)/d alpha_c = relu'(alpha_c) / w exactly (no shared term)
    w = float(torch.relu(m.alpha).sum() + torch.relu(m.beta).sum())
    np.testing.assert_allclose(m.alpha.grad.flatten().numpy(), 3.0 / w, rtol=1e-5)
