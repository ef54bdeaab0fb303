"""Pin the CPU oracle (oracle/reference_oracle.py) to vectors produced by the UNMODIFIED reference
modules (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from nerf_qa_b200 import synth
from oracle import reference_oracle as O

DISTS_CASES = ["dists_64", "dists_odd_45x77", "dists_135x240", "dists_256"]
ADISTS_CASES = ["adists_96x128", "adists_176", "adists_256", "adists_352x368"]


def _inputs(g):
    b, h, w, seed = (int(v) for v in g["shape"])
    return synth.make_frame_pairs(b, h, w, seed=seed)


@pytest.mark.parametrize("name", DISTS_CASES)
def test_dists_oracle_matches_reference(name, golden, vgg_weights, alpha_beta):
    g = golden(name)
    x, y = _inputs(g)
    # the synthetic generator itself must be reproducible across machines
    np.testing.assert_allclose([float(x.double().sum()), float(y.double().sum())], g["input_checksum"][:2], rtol=1e-9)
    ws, bs = vgg_weights
    a, b = alpha_beta
    with torch.no_grad():
        fx, fy = O.trunk(x, ws, bs), O.trunk(y, ws, bs)
        score = O.dists_from_feats(fx, fy, a, b)
    assert [list(f.shape) for f in fx] == g["level_shapes"].tolist()
    np.testing.assert_allclose(fx[5][0, :8].numpy(), g["relu5_3_x0"], rtol=1e-5, atol=1e-6)
    s1 = torch.cat([O.dists_level_stats(fx[k], fy[k])[0] for k in range(6)], 1).numpy()
    s2 = torch.cat([O.dists_level_stats(fx[k], fy[k])[1] for k in range(6)], 1).numpy()
    np.testing.assert_allclose(s1, g["s1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(s2, g["s2"], rtol=1e-4, atol=1e-5)
    # fp32 score equals the reference's fp32 score to rounding, and sits within 1e-4 rel of its fp64 score
    np.testing.assert_allclose(score.numpy(), g["score_f32"], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(score.numpy(), g["score_f64"], rtol=1e-4)
    avg = O.dists_from_feats(fx, fy, a, b, batch_average=True)
    np.testing.assert_allclose(float(avg), float(g["score_batch_avg"]), rtol=2e-6)


@pytest.mark.parametrize("name", ADISTS_CASES)
def test_adists_oracle_matches_reference(name, golden, vgg_weights):
    g = golden(name)
    x, y = _inputs(g)
    ws, bs = vgg_weights
    with torch.no_grad():
        fx, fy = O.trunk(x, ws, bs), O.trunk(y, ws, bs)
        score, ps, wts, _ = O.adists_from_feats(fx, fy, return_parts=True)
        swapped = O.adists_from_feats(fy, fx)
    assert [list(p.shape) for p in ps] == g["ps_shapes"].tolist()
    summ = np.stack([np.array([float(p.double().mean()), float(p.double().std()), float(p.min()), float(p.max())]) for p in ps])
    np.testing.assert_allclose(summ, g["ps_summary"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ps[3].numpy(), g["ps_coarse"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ps[5].numpy(), g["ps_level5"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(score.numpy(), g["score_f32"], rtol=5e-6, atol=5e-7)
    np.testing.assert_allclose(score.numpy(), g["score_f64"], rtol=1e-4)
    np.testing.assert_allclose(swapped.numpy(), g["score_swapped"], rtol=5e-6, atol=5e-7)
    np.testing.assert_allclose(float(1 - (1 - score).mean()), float(g["score_as_loss"]), rtol=5e-6)
    assert abs(float(wts.sum(1).mean()) - 1.0) < 1e-5


def test_l2pool_filter_is_3x3_hanning():
    f = O.hanning_filter()
    np.testing.assert_allclose(f.numpy(), np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float32) / 16, rtol=1e-6)
    x = torch.rand(1, 2, 5, 7)
    assert list(O.l2pool(x).shape) == [1, 2, 3, 4]


def test_project_weights(alpha_beta):
    a, b = alpha_beta
    pa, pb = O.project_weights(a.clone() - 0.001, b.clone())
    assert abs(float(pa.sum() + pb.sum()) - 1) < 1e-5
    assert float(pa[:3].min()) > 0.019 and float(pb.min()) >= 0


def test_pooling_and_correlations():
    r = np.random.RandomState(0)
    s = r.rand(17)
    p = O.pool_video(s)
    assert abs(p["mean"] - s.mean()) < 1e-12 and p["min"] == s.min() and p["max"] == s.max()
    a = r.rand(20); b = a + 0.1 * r.rand(20)
    c = O.correlations(a, b)
    assert 0.5 < c["srcc"] <= 1 and 0.5 < c["plcc"] <= 1 and 0.3 < c["ktcc"] <= 1


def test_oracle_matches_reference_at_1080p(golden, vgg_weights, alpha_beta):
    """The benchmarked shape (BASELINE configs[1] / [2]): one 1920x1080 pair, every A-DISTS level windowed.  The fixture
    is what the unmodified reference DISTS / ADISTS modules return (tests/golden/make_golden_1080p.py)."""
    g = golden("pair_1080p")
    x, y = _inputs(g)
    np.testing.assert_allclose([float(x.double().sum()), float(y.double().sum())], g["input_checksum"], rtol=1e-9)
    ws, bs = vgg_weights
    a, b = alpha_beta
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        fx, fy = O.trunk(x, ws, bs), O.trunk(y, ws, bs)
        sd = O.dists_from_feats(fx, fy, a, b)
        sa, ps, wts, terms = O.adists_from_feats(fx, fy, return_parts=True)
    np.testing.assert_allclose(sd.numpy(), g["dists_f32"], rtol=5e-5, atol=2e-7)
    np.testing.assert_allclose(sa.numpy(), g["adists_f32"], rtol=1e-4, atol=5e-7)
    np.testing.assert_allclose(torch.stack(terms, 1).numpy(), g["adists_level_terms"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(wts.squeeze(-1).numpy(), g["adists_weights"], rtol=1e-5, atol=1e-9)
    assert [list(p.shape) for p in ps] == g["ps_shapes"].tolist()
    np.testing.assert_allclose(ps[5].numpy(), g["ps_level5"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ps[3].numpy(), g["ps_level3"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ps[1][..., ::8, ::8].numpy(), g["ps_level1_dec8"], rtol=1e-4, atol=1e-5)
