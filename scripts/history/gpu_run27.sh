mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -k "adists or shared_trunk" > gpurun_out/t_27.log 2>&1; echo "rc=$?"
tail -4 gpurun_out/t_27.log
timeout 600 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1s_adists.json 2> gpurun_out/bench_r1s_adists.err; echo "adists rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_r1s_adists.json').read().strip().splitlines()[-1]);print(d['value'], {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1s_adists.csv python bench.py --workload adists --steps 1 --warmup 3 --pairs 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1s_adists.log 2>&1; echo "ncu rc=$?"
