#!/bin/bash
# backward: tensor-pipe conv1_1 backward + rewritten level_grad
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q -k "grad or backward or flavour or frmodel" > gpurun_out/r39_test.log 2>&1
echo "test exit $?" >> gpurun_out/r39_test.log
tail -4 gpurun_out/r39_test.log
for tc in 1 0; do
NQA_CONV11_BWD_TC=$tc timeout 300 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r39_bench_train_tc$tc.log 2>&1
echo "conv11 bwd tc=$tc"; grep -o "\"value\": [0-9.]*\|\"kernel_families\": {.*}}" gpurun_out/r39_bench_train_tc$tc.log | head -3
done
