mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gradient or training_step or flavours" -s > gpurun_out/t_33.log 2>&1; echo "bwd rc=$?"
grep -h "parity\] d score\|512x512\|passed\|failed\|rror" gpurun_out/t_33.log | tail -14
timeout 600 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_r1y_train.json 2> gpurun_out/bench_r1y_train.err; echo "train rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_r1y_train.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print(round(d['value'],1), d['ms_per_step'], {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
