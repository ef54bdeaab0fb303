#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3o}
timeout 600 python scripts/r3_one_term_sweep.py > gpurun_out/${TAG}_one_term.log 2>&1; tail -8 gpurun_out/${TAG}_one_term.log
for thr in 400000 100000 30000 0; do
NQA_ADISTS_ONE_TERM_MIN_PIXELS=$thr timeout 600 python bench.py --workload adists --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_thr$thr.json 2> /dev/null
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench_thr$thr.json').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print($thr, round(d['value'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
done
