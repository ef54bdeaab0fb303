mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all3.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/t_gpu_all3.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo "bench rc=$?"
cat gpurun_out/bench_r1e.json
NQA_CONV_N64=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1e_n128.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1e.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stage1_fused|pool_stats|halo" -c 7 -o gpurun_out/prof_r1e python bench.py --steps 1 --warmup 3 --pairs 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_r1e.log 2>&1; echo "ncu full rc=$?"
