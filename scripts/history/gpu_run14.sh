mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -x -q -m gpu -k "adists" > gpurun_out/t_adists14.log 2>&1; echo "adists tests rc=$?"
tail -3 gpurun_out/t_adists14.log
timeout 600 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1o_adists.json 2> gpurun_out/bench_r1o_adists.err; echo "adists rc=$?"
cut -c1-260 gpurun_out/bench_r1o_adists.json; python -c "
import json;d=json.loads(open('gpurun_out/bench_r1o_adists.json').read().strip().splitlines()[-1]);print({k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})"
