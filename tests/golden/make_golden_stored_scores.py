"""Fixture of the reference's own STORED results for the video-level consumers of the hot path (SURVEY.md section 8,
row f3): run in the build container only (/root/reference is absent on the GPU box).

    python tests/golden/make_golden_stored_scores.py   ->   tests/golden/stored_scores.npz

From /root/reference/scores_aspect.csv (written by test2_prep.py:183-193,512 -- 56 Test-2 videos scored by the reference
DISTS / ADISTS modules with the pretrained VGG16): the per-video mean / std / min / max of both metrics, MOS, the
reference video (= scene) of every render, and the PER-FRAME scores recovered from the stored `frame_bias_*` vectors
(test2_prep.py:167-168: bias = video mean - frame score, printed with 7 significant digits).
From /root/reference/results_combined.csv (prep.py:168-232): DISTS, MOS, DMOS and scene of the 48 NeRF-QA videos, and
from /root/reference/correlations.csv row `DISTS` (prep.py:394) the correlations the reference computed from them.
"""
import ast
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    t = pd.read_csv(os.path.join(REF, "scores_aspect.csv"))
    scenes = sorted(t["reference_folder"].unique())
    out = {
        "test2_scene_names": np.array(scenes),
        "test2_scene": t["reference_folder"].map({s: i for i, s in enumerate(scenes)}).to_numpy(np.int64),
        "test2_mos": t["MOS"].to_numpy(np.float64),
        "test2_frame_count": t["frame_count"].to_numpy(np.int64),
    }
    for metric, col, bias in (("dists", "DISTS", "frame_bias_dists"), ("adists", "A-DISTS", "frame_bias_adists")):
        frames = []
        for mean, b in zip(t[col], t[bias]):
            frames.append(np.float64(mean) - np.array([float(v) for v in ast.literal_eval(b)], dtype=np.float64))
        assert [len(f) for f in frames] == t["frame_count"].tolist()
        out[f"test2_{metric}_frames"] = np.concatenate(frames)
        for stat, suffix in (("mean", ""), ("std", "_std"), ("min", "_min"), ("max", "_max")):
            out[f"test2_{metric}_{stat}"] = t[col + suffix].to_numpy(np.float64)
    n = pd.read_csv(os.path.join(REF, "results_combined.csv"))
    nscenes = sorted(n["scene"].unique())
    out.update(nerfqa_dists=n["DISTS"].to_numpy(np.float64), nerfqa_mos=n["MOS"].to_numpy(np.float64),
               nerfqa_dmos=n["DMOS"].to_numpy(np.float64), nerfqa_scene_names=np.array(nscenes),
               nerfqa_scene=n["scene"].map({s: i for i, s in enumerate(nscenes)}).to_numpy(np.int64))
    c = pd.read_csv(os.path.join(REF, "correlations.csv")).set_index("Metric").loc["DISTS"]
    out["nerfqa_corr_names"] = np.array(list(c.index))
    out["nerfqa_corr_values"] = c.to_numpy(np.float64)
    np.savez_compressed(os.path.join(HERE, "stored_scores.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
