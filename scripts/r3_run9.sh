#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3v}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_t_gpu.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${TAG}_t_gpu.log
run() { name=$1; shift; timeout 900 python bench.py "$@" --no-cpu-baseline > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "$name rc=$?"; }
run dists_bf16x3 --workload dists --precision bf16x3 --steps 3 --warmup 3 --no-e2e --headline-only
run train_bf16x3 --workload train --precision bf16x3 --steps 5 --warmup 3
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*.json')):
    l=[x for x in open(f).read().splitlines() if x.startswith('{')]
    if not l: print(f,'NO LINE'); continue
    d=json.loads(l[-1]); r=d.get('roofline') or {}
    print(f.split('bench_')[1][:-5].ljust(14), round(d['value'],1), 'top', r.get('family'), 'frac', r.get('frac'), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
