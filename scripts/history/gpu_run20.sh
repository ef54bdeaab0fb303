mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/t_gpu_all20.log 2>&1; echo "all rc=$?"
grep -h "flat\|A-DISTS (\|DISTS (1, 1\|DISTS (2, 3\|passed\|failed\|Error" gpurun_out/t_gpu_all20.log | tail -30
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1q.json 2> gpurun_out/bench_r1q.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r1q.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})"
