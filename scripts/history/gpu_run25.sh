mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -k "adists or shared_trunk or smoke" > gpurun_out/t_25.log 2>&1; echo "rc=$?"
tail -4 gpurun_out/t_25.log
timeout 600 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1r_adists.json 2> gpurun_out/bench_r1r_adists.err; echo "adists rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r1r_adists.json').read().strip().splitlines()[-1]);print(d['value'], {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})"
