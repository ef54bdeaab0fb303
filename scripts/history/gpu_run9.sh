mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gradient or training_step" -s > gpurun_out/t_bwd.log 2>&1; echo "bwd rc=$?"
grep -h "parity\|passed\|failed\|Error\|error" gpurun_out/t_bwd.log | tail -30
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all9.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/t_gpu_all9.log
timeout 600 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1l_adists.json 2> gpurun_out/bench_r1l_adists.err; echo "adists rc=$?"
cut -c1-200 gpurun_out/bench_r1l_adists.json
