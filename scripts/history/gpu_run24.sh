mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "shared_trunk" -s > gpurun_out/t_24.log 2>&1; echo "rc=$?"
grep -h "parity\|passed\|failed\|rror" gpurun_out/t_24.log | tail -8
python - <<'PY' > gpurun_out/shared_trunk_timing.log 2>&1
import torch, time, warnings, sys
sys.path.insert(0, '.')
warnings.simplefilter('ignore')
from tests._util import make_dists, make_adists
from nerf_qa_b200 import synth
sys.argv=['x']
import bench
d, a = make_dists("bf16"), make_adists("bf16")
x, y = bench.make_frames_device(16, torch.device('cuda:0'))
def t(f, n=3):
    for _ in range(2): f()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n
with torch.no_grad():
    sep = t(lambda: (d(x, y), a(x, y, as_loss=False)))
    both = t(lambda: d.score_with_adists(x, y))
print(f"16 pairs 1080p: separate DISTS + A-DISTS {sep*1e3:.1f} ms ({16/sep:.1f} pairs/s), shared trunk {both*1e3:.1f} ms ({16/both:.1f} pairs/s)")
PY
cat gpurun_out/shared_trunk_timing.log | tail -3
