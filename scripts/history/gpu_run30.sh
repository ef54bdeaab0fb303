mkdir -p gpurun_out
set -x
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
NQA_LANES=1 timeout 300 $B > gpurun_out/bench_r1w_l1.json 2>&1
NQA_LANES=2 timeout 300 $B > gpurun_out/bench_r1w_l2.json 2>&1
NQA_LANES=1 timeout 300 $B > gpurun_out/bench_r1w_l1b.json 2>&1
NQA_LANES=2 timeout 300 $B > gpurun_out/bench_r1w_l2b.json 2>&1
python - <<'PY'
import json
for f in ('l1','l2','l1b','l2b'):
    l=[x for x in open(f'gpurun_out/bench_r1w_{f}.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print(f, round(d['value'],1), d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_30.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_30.log
