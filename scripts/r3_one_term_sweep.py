"""A-DISTS window kernel: how many pixels a map needs before its operands go into the MMAs as ONE bf16 term (rounded
products, rounded first-pass results) instead of two (exact hi + lo).  Scores of the 1080p golden pair for a range of
thresholds against the all-two-terms scores (same features: isolates the window arithmetic) and the reference golden,
plus the A-DISTS throughput of each setting."""
import json, os, subprocess, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._util import make_adists, golden_inputs

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pair_1080p.npz"))
x, y = golden_inputs(g)
xs, ys = x.cuda(), y.cuda()
out = {}
for thr in ("-1", "400000", "100000", "30000", "5000", "0"):
    if thr == "-1":
        os.environ["NQA_ADISTS_MMA_TERMS"] = "2"
    else:
        os.environ.pop("NQA_ADISTS_MMA_TERMS", None)
        os.environ["NQA_ADISTS_ONE_TERM_MIN_PIXELS"] = thr
    with torch.no_grad():
        m = make_adists("bf16")
        s = m(xs, ys, as_loss=False).cpu().numpy().astype(np.float64)
    out[thr] = s
    del m
ref = out["-1"]
for thr, s in out.items():
    print(f"one-term >= {thr:>7} px: score {s}, vs two-term everywhere {np.max(np.abs(s - ref)):.3e}, vs reference fp32 golden "
          f"{np.max(np.abs(s - g['adists_f32'])):.3e}")
