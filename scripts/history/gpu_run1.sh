mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused" -s > gpurun_out/t_fused.log 2>&1; echo "fused rc=$?"
tail -15 gpurun_out/t_fused.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all2.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/t_gpu_all2.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench rc=$?"
cat gpurun_out/bench_r1d.json
NQA_STAGE1_FUSED=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1d_nos1.json 2>&1
NQA_POOL_STATS_FUSED=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1d_nops.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1d.log 2>&1; echo "ncu rc=$?"
