#!/bin/bash
# Full GPU suite + the both / video workloads.   gpurun -- 'bash scripts/r3_run5.sh TAG'
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3h}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_t_gpu.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/${TAG}_t_gpu.log
for wl in both adists; do
timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err; echo "bench $wl rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench_$wl.json').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(round(d['value'],1), 'e2e', (round(d['e2e']['value'],1) if d.get('e2e') else None), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
done
