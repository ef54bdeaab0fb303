mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all4.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/t_gpu_all4.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench rc=$?"
cat gpurun_out/bench_r1f.json
NQA_CONV_N64=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1f_n128.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1f.log 2>&1; echo "ncu rc=$?"
