#!/bin/bash
# Round 2, run 5: the other workloads / precisions quoted in DESIGN.md section 6.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2u}
run() { name=$1; shift; timeout 600 python bench.py "$@" --no-cpu-baseline > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "$name rc=$?"; tail -1 gpurun_out/${TAG}_bench_${name}.err; }
run both --workload both --steps 5 --warmup 3
run dists_bf16x3 --workload dists --precision bf16x3 --steps 3 --warmup 3 --no-e2e
run dists_fp32 --workload dists --precision fp32 --pairs 16 --steps 2 --warmup 3 --no-e2e
run adists_fp32 --workload adists --precision fp32 --pairs 8 --steps 2 --warmup 3 --no-e2e
run train_bf16 --workload train --steps 5 --warmup 3
run train_bf16x3 --workload train --precision bf16x3 --steps 5 --warmup 3
run train_fp32 --workload train --precision fp32 --steps 2 --warmup 3
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*.json')):
    l=[x for x in open(f).read().splitlines() if x.startswith('{')]
    if not l: print(f,'NO LINE'); continue
    d=json.loads(l[-1]); r=d['roofline']
    print(f.split('bench_')[1][:-5].ljust(14), round(d['value'],1), 'e2e', (round(d['e2e']['value'],1) if d.get('e2e') else None), 'top', r['family'], 'frac', round(r['frac'],3) if r['frac'] else None, {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
