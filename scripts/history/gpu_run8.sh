mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all8.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/t_gpu_all8.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; echo "bench rc=$?"
cat gpurun_out/bench_r1k.json
timeout 600 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1k_adists.json 2> gpurun_out/bench_r1k_adists.err; echo "adists rc=$?"
cat gpurun_out/bench_r1k_adists.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1k_adists.csv python bench.py --workload adists --steps 1 --warmup 3 --pairs 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1k_adists.log 2>&1; echo "ncu rc=$?"
