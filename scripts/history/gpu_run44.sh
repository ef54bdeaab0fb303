#!/bin/bash
# plain remote arrives: conv tests + stage-1 bit identity, bench single vs pair stage 1
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q -k "fused_stage1 or conv or golden or 1080p" > gpurun_out/r44_test.log 2>&1
echo "test exit $?" >> gpurun_out/r44_test.log
tail -4 gpurun_out/r44_test.log
for v in "0" "1"; do
  NQA_STAGE1_PAIR=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r44_bench_p$v.log 2>&1
  echo "stage1 pair=$v"
  grep -o "\"value\": [0-9.]*\|\"stage1\": {[^}]*}\|\"conv_tc\": {[^}]*}\|sm_mhz\": [0-9.]*" gpurun_out/r44_bench_p$v.log
done
