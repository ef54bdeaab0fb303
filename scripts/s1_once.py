"""One 8-pair 1080p DISTS call (for ncu captures of single kernels)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_qa_b200 import DISTS, synth
m = DISTS(precision="bf16").cuda()
x, y = synth.make_frame_pairs(1, 1080, 1920, seed=3)
x = x.cuda().repeat(8, 1, 1, 1).contiguous(); y = y.cuda().repeat(8, 1, 1, 1).contiguous()
with torch.no_grad():
    for _ in range(2):
        s = m(x, y)
torch.cuda.synchronize()
print(s[:2].tolist())
