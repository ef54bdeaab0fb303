"""One 2-pair 1080p A-DISTS call (for ncu captures of single kernels)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_qa_b200 import ADISTS, synth
m = ADISTS(precision="bf16").cuda()
x, y = synth.make_frame_pairs(2, 1080, 1920, seed=3)
x = x.cuda(); y = y.cuda()
with torch.no_grad():
    for _ in range(2):
        s = m(x, y)
torch.cuda.synchronize()
print(s.tolist())
