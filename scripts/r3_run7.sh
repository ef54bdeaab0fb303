#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3r}
timeout 900 python -m pytest tests -x -q -m gpu -k "pool or dists_bf16 or batch_independent or 1080p or shared_trunk or training or host_path" > gpurun_out/${TAG}_t.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${TAG}_t.log
timeout 600 python bench.py --headline-only --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench.json').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(round(d['value'],1), 'e2e', (round(d['e2e']['value'],1) if d.get('e2e') else None), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
