mkdir -p gpurun_out
set -x
timeout 600 python bench.py --precision bf16x3 --pairs 16 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1z_x3.json 2>&1
timeout 600 python bench.py --precision fp32 --pairs 8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1z_fp32.json 2>&1
python - <<'PY'
import json
for f in ('x3','fp32'):
    l=[x for x in open(f'gpurun_out/bench_r1z_{f}.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print(f, round(d['value'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
