"""Row f3: the video-level consumers of the per-frame scores -- temporal pooling and the
PLCC / SRCC / KTCC tables -- reproduce the numbers the reference itself stored (tests/golden/stored_scores.npz, made by
tests/golden/make_golden_stored_scores.py from /root/reference/scores_aspect.csv, results_combined.csv and
correlations.csv).  "Video-level SRCC / PLCC unchanged" (north_star) is checked on these real per-frame scores with a
bf16-sized perturbation."""
import numpy as np
import pytest

from nerf_qa_b200 import video


@pytest.fixture(scope="module")
def stored(golden):
    return golden("stored_scores")


def test_temporal_pooling_reproduces_the_stored_video_statistics(stored):
    """prep.py:191-198 / test2_prep.py:158-165: np.mean / np.std / np.min / np.max over a video's frames.  The frames come
    from the stored frame_bias vectors (7 significant digits), hence the tolerances."""
    lengths = stored["test2_frame_count"]
    assert int(lengths.sum()) == stored["test2_dists_frames"].size == 6562
    for metric in ("dists", "adists"):
        pooled = video.pool_videos(stored[f"test2_{metric}_frames"], lengths)
        np.testing.assert_allclose(pooled["mean"], stored[f"test2_{metric}_mean"], rtol=0, atol=2e-7)
        np.testing.assert_allclose(pooled["std"], stored[f"test2_{metric}_std"], rtol=2e-5, atol=2e-7)
        np.testing.assert_allclose(pooled["min"], stored[f"test2_{metric}_min"], rtol=0, atol=2e-7)
        np.testing.assert_allclose(pooled["max"], stored[f"test2_{metric}_max"], rtol=0, atol=2e-7)


def test_correlations_reproduce_the_stored_tables(stored):
    """stored Test-2 DISTS / A-DISTS vs MOS, and correlations.csv row 10 (NeRF-QA DISTS)."""
    want = {"dists": (0.8557, 0.8643, 0.6866), "adists": (0.8133, 0.8199, 0.6281)}
    lengths = stored["test2_frame_count"]
    for metric, (plcc, srcc, ktcc) in want.items():
        pooled = video.pool_videos(stored[f"test2_{metric}_frames"], lengths)
        c = video.correlations(pooled["mean"], stored["test2_mos"])
        assert abs(abs(c["plcc"]) - plcc) < 5e-5 and abs(abs(c["srcc"]) - srcc) < 5e-5 and abs(abs(c["ktcc"]) - ktcc) < 5e-5
        c2 = video.correlations(stored[f"test2_{metric}_mean"], stored["test2_mos"])
        assert abs(c["srcc"] - c2["srcc"]) < 1e-12 and abs(c["ktcc"] - c2["ktcc"]) < 1e-12
    c = video.correlations(stored["nerfqa_dists"], stored["nerfqa_mos"])
    names = list(stored["nerfqa_corr_names"])
    ref = dict(zip(names, stored["nerfqa_corr_values"]))
    assert abs(abs(c["plcc"]) - ref["all mos plcc"]) < 1e-12          # correlations.csv:10 = 0.78967 / 0.76769 / 0.56383
    assert abs(abs(c["srcc"]) - ref["all mos srcc"]) < 1e-12
    assert abs(abs(c["ktcc"]) - ref["all mos ktcc"]) < 1e-12
    d = video.correlations(stored["nerfqa_dists"], stored["nerfqa_dmos"])
    assert abs(abs(d["plcc"]) - ref["all dmos plcc"]) < 1e-12 and abs(abs(d["srcc"]) - ref["all dmos srcc"]) < 1e-12


def test_summary_matches_the_logger_on_the_stored_scores(stored):
    """video.summarize = MetricCollectionLogger.log_summary without wandb (logger.py:144-282): per-video means, per-scene
    means, real / synthetic splits.  NeRF-QA: one 'frame' per video (the CSV holds video-level scores); the synthetic /
    real split reproduces the `syn` / `tnt` columns of correlations.csv:10."""
    names = stored["nerfqa_scene_names"]
    scene = names[stored["nerfqa_scene"]]
    n = len(scene)
    s = video.summarize(stored["nerfqa_dists"], np.arange(n), scene, mos=stored["nerfqa_mos"])
    ref = dict(zip(list(stored["nerfqa_corr_names"]), stored["nerfqa_corr_values"]))
    # prep.py:225-228 calls ficus / ship / drums `syn` (lego left out) and the four Tanks-and-Temples scenes `tnt`; the
    # logger's synthetic list (logger.py:218) includes lego
    syn = np.isin(scene, ["ficus", "ship", "drums"])
    c_syn = video.correlations(stored["nerfqa_dists"][syn], stored["nerfqa_mos"][syn])
    lego_too = np.isin(scene, list(video.SYNTH_SCENES))
    c_synth = video.correlations(stored["nerfqa_dists"][lego_too], stored["nerfqa_mos"][lego_too])
    for k in ("plcc", "srcc", "ktcc"):
        assert abs(abs(s["correlations"][k]) - ref[f"all mos {k}"]) < 1e-12
        assert abs(abs(c_syn[k]) - ref[f"syn mos {k}"]) < 1e-12
        assert abs(abs(s["real_correlations"][k]) - ref[f"tnt mos {k}"]) < 1e-12
        assert abs(s["synthetic_correlations"][k] - c_synth[k]) < 1e-12
    assert len(s["scenes"]) == 8 and len(s["scene_correlations"]) == 8
    for i, sc in enumerate(s["scenes"]):
        np.testing.assert_allclose(s["scene_mean"][i], stored["nerfqa_dists"][scene == sc].mean(), rtol=1e-12)
    assert abs(s["average_over_scenes"] - np.mean(s["scene_mean"])) < 1e-15
    assert 0 <= s["scene_min_srcc"] <= s["scene_mean_srcc"] <= 1

    # Test-2 with real per-frame scores: video ids per frame, scene = the reference video of the render
    lengths = stored["test2_frame_count"]
    vid, _ = video.frame_to_video(lengths)
    scene_f = stored["test2_scene_names"][stored["test2_scene"]][vid]
    mos_f = stored["test2_mos"][vid]
    t = video.summarize(stored["test2_dists_frames"], vid, scene_f, mos=mos_f)
    np.testing.assert_allclose(t["video_mean"], stored["test2_dists_mean"], atol=2e-7)
    # like the reference (logger.py:160) the per-video MOS is np.mean over the video's frames: equal MOS values of two
    # videos can come out one ulp apart and break a tie, so the rank correlations move in the 4th decimal
    assert abs(abs(t["correlations"]["srcc"]) - 0.8643) < 1e-3 and abs(abs(t["correlations"]["plcc"]) - 0.8557) < 5e-5
    per_scene = [stored["test2_dists_mean"][stored["test2_scene"] == i].mean() for i in range(len(stored["test2_scene_names"]))]
    np.testing.assert_allclose(t["scene_mean"], per_scene, atol=2e-7)


def test_video_level_correlations_survive_a_bf16_sized_perturbation(stored):
    """north_star: "video-level SRCC/PLCC against the stored scores must be unchanged".  The bf16 trunk moves a frame score
    by at most 2e-3 absolute (measured <= 1.3e-3, mostly a per-video bias): perturbing the reference's real per-frame
    scores by that much leaves the Test-2 correlations where they are to two decimals."""
    r = np.random.RandomState(0)
    lengths = stored["test2_frame_count"]
    vid, _ = video.frame_to_video(lengths)
    for metric in ("dists", "adists"):
        frames = stored[f"test2_{metric}_frames"]
        base = video.correlations(video.pool_videos(frames, lengths)["mean"], stored["test2_mos"])
        worst = 0.0
        for _ in range(20):
            noisy = frames + r.uniform(-1.3e-3, 1.3e-3, size=len(lengths))[vid] + r.uniform(-5e-4, 5e-4, size=frames.size)
            c = video.correlations(video.pool_videos(noisy, lengths)["mean"], stored["test2_mos"])
            worst = max(worst, abs(c["plcc"] - base["plcc"]), abs(c["srcc"] - base["srcc"]))
        assert worst < 1e-2, worst
