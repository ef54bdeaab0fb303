#!/bin/bash
# A-DISTS after the branch-free window loads; per-kernel launch lists of the train and A-DISTS workloads
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q -k "adists or as_map or window" > gpurun_out/r38_test.log 2>&1
echo "test exit $?" >> gpurun_out/r38_test.log
tail -3 gpurun_out/r38_test.log
timeout 300 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r38_bench_adists.log 2>&1
grep -o "\"value\": [0-9.]*\|\"kernel_families\": {.*}}" gpurun_out/r38_bench_adists.log | head -3
timeout 300 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r38_bench_train.log 2>&1
grep -o "\"value\": [0-9.]*\|\"kernel_families\": {.*}}" gpurun_out/r38_bench_train.log | head -3
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r38_launches_train.csv python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r38_ncu_train.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r38_launches_adists.csv python bench.py --workload adists --pairs 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r38_ncu_adists.log 2>&1
ls -la gpurun_out | tail -8
