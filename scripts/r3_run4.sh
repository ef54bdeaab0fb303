#!/bin/bash
# A-DISTS subset + A-DISTS bench (+ optional kernel profile):  gpurun -- 'bash scripts/r3_run4.sh TAG [REGEX]'
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3f}
timeout 900 python -m pytest tests -x -q -m gpu -k "adists or small or shared_trunk" > gpurun_out/${TAG}_t_adists.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${TAG}_t_adists.log
timeout 600 python bench.py --workload adists --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_adists.json 2> gpurun_out/${TAG}_bench_adists.err; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench_adists.json').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(round(d['value'],1), 'e2e', (round(d['e2e']['value'],1) if d.get('e2e') else None), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
if [ -n "$2" ]; then bash scripts/r3_prof_kernel.sh $TAG "$2" ${3:-2}; fi
