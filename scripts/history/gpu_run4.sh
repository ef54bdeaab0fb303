mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all5.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/t_gpu_all5.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo "bench rc=$?"
cat gpurun_out/bench_r1g.json
NQA_CONV_HALO_MAXCIN=256 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1g_h256.json 2>&1
NQA_CONV_HALO_MAXCIN=512 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1g_h512.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1g.log 2>&1; echo "ncu rc=$?"
NQA_CONV_HALO_MAXCIN=512 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g_h512.csv python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1g_h512.log 2>&1; echo "ncu rc=$?"
