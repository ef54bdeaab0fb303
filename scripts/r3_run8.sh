#!/bin/bash
# bf16x3 on the CTA-pair kernel: the bf16x3 tests + the DISTS / training benches of that precision.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3t}
timeout 900 python -m pytest tests -x -q -m gpu -k "bf16x3 or features_per_level or frame_gradients or edge_shapes or conv3x3 or fused_pool or batch_independent or large_frames or host_path" -s > gpurun_out/${TAG}_t.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|Error" gpurun_out/${TAG}_t.log | tail -3; grep -E "parity\].*bf16x3" gpurun_out/${TAG}_t.log | head -12
for ps in 1 0; do
NQA_PAIR_SPLIT=$ps timeout 600 python bench.py --workload dists --precision bf16x3 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --headline-only > gpurun_out/${TAG}_bench_x3_ps$ps.json 2> gpurun_out/${TAG}_bench_x3_ps$ps.err; echo "bench ps=$ps rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench_x3_ps$ps.json').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('pair_split=$ps', round(d['value'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
done
