mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/t_gpu_all16.log 2>&1; echo "all rc=$?"
grep -h "as_map\|passed\|failed\|Error" gpurun_out/t_gpu_all16.log | tail -10
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1p.json 2> gpurun_out/bench_r1p.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r1p.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e']['value'], d['clocks'])"
