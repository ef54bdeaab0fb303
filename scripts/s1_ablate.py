"""Stage-1 ablation timing (temporary debug switches NQA_S1_DBG in stage1_fused.cu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_qa_b200 import DISTS, synth
m = DISTS(precision="bf16").cuda()
x, y = synth.make_frame_pairs(1, 1080, 1920, seed=3)
x = x.cuda().repeat(8, 1, 1, 1).contiguous(); y = y.cuda().repeat(8, 1, 1, 1).contiguous()
ctx = m._context(x.device)
for d in [0, 1, 2, 4, 8, 16, 32, 64, 128, 2 | 8, 4 | 64 | 128 | 32, 1 | 2 | 8 | 16, 255]:
    os.environ["NQA_S1_DBG"] = str(d)
    with torch.no_grad():
        for _ in range(2):
            m(x, y)
        ctx.profile_read(); ctx.profile_enable(True)
        for _ in range(3):
            m(x, y)
        ctx.profile_enable(False)
    ms, n = ctx.profile_read()["stage1"]
    print(f"dbg={d:3d} stage1 {ms / n:.3f} ms/launch ({n} launches)", flush=True)
