#!/bin/bash
# full GPU suite + smoke + the benches after the backward / tape / load-path changes
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r41_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r41_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r41_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r41_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r41_bench.json 2> gpurun_out/r41_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload adists --steps 3 --warmup 3 > gpurun_out/r41_bench_adists.json 2> gpurun_out/r41_bench_adists.err; echo "adists rc=$?"
timeout 600 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/r41_bench_train.json 2> gpurun_out/r41_bench_train.err; echo "train rc=$?"
timeout 600 python bench.py --precision bf16x3 --pairs 16 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r41_bench_x3.json 2>&1
python - <<'PY'
import json
for f in ('r41_bench','r41_bench_adists','r41_bench_train','r41_bench_x3'):
    l=[x for x in open(f'gpurun_out/{f}.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print(f, d['metric'], round(d['value'],2), 'e2e', (round(d['e2e']['value'],2) if d.get('e2e') else None), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
