#!/bin/bash
# Round-end sanity: full GPU suite, smoke, and the bench workloads (gpurun -- 'bash scripts/gpu_final.sh').
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests.log
grep "conv1_1 backward tensor pipe" gpurun_out/final_tests.log | head -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_adists.json 2> gpurun_out/final_bench_adists.err; echo "adists rc=$?"
timeout 600 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_train.json 2> gpurun_out/final_bench_train.err; echo "train rc=$?"
python - <<'PY'
import json
for f in ('final_bench','final_bench_adists','final_bench_train'):
    l=[x for x in open(f'gpurun_out/{f}.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print(f, d['metric'], round(d['value'],2), 'e2e', (round(d['e2e']['value'],2) if d.get('e2e') else None), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
