#!/bin/bash
# activation tape: tests + train bench with / without
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q -k "grad or backward or flavour or frmodel or tape or saved" > gpurun_out/r40_test.log 2>&1
echo "test exit $?" >> gpurun_out/r40_test.log
tail -15 gpurun_out/r40_test.log
for mb in ; do
NQA_TAPE_MAX_MB=$mb timeout 300 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r40_bench_train_tape$mb.log 2>&1
echo "tape limit=$mb"; grep -o "\"value\": [0-9.]*\|\"kernel_families\": {.*}}" gpurun_out/r40_bench_train_tape$mb.log | head -3
done
