"""One profiled call of a scoring workload for ncu (`ncu --profile-from-start off`): two warm-up calls, then the third
between cudaProfilerStart / Stop.   python scripts/profile_once.py {dists,adists,both} [pairs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_qa_b200 import synth
from tests._util import make_adists, make_dists

what = sys.argv[1] if len(sys.argv) > 1 else "dists"
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
x, y = synth.make_frame_pairs(pairs, 1080, 1920, seed=3)
x, y = x.cuda(), y.cuda()
d, a = make_dists("bf16"), make_adists("bf16")
call = {"dists": lambda: d(x, y), "adists": lambda: a(x, y, as_loss=False), "both": lambda: d.score_with_adists(x, y)[1]}[what]
with torch.no_grad():
    for _ in range(2):
        s = call()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    s = call()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print(what, s[:2].tolist())
