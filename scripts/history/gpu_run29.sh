mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "tcgen05_bf16 and not bf16x3" > gpurun_out/t_29.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/t_29.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/bench_r1u_base.json 2>&1
NQA_PAIR_MINCIN=64 timeout 300 $B > gpurun_out/bench_r1u_p64.json 2>&1
python - <<'PY'
import json
for f in ('base','p64'):
    l=[x for x in open(f'gpurun_out/bench_r1u_{f}.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print(f, round(d['value'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
