mkdir -p gpurun_out
set -x
timeout 180 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "tcgen05_bf16 and not bf16x3" -s > gpurun_out/t_pair.log 2>&1; rc=$?; echo "pair ops rc=$rc"
tail -20 gpurun_out/t_pair.log
if [ $rc -ne 0 ]; then nvidia-smi > gpurun_out/smi_after_fail.log 2>&1; exit 1; fi
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all7.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/t_gpu_all7.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; echo "bench rc=$?"
cat gpurun_out/bench_r1i.json
NQA_CONV_PAIR=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r1i_nopair.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1i.log 2>&1; echo "ncu rc=$?"
