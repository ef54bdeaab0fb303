#!/bin/bash
# fused stage 1 variants: bit-identity test, then bench of each
cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_stage1" > gpurun_out/r35_test.log 2>&1
echo "test exit $?" >> gpurun_out/r35_test.log
tail -5 gpurun_out/r35_test.log
for v in "0 1" "0 0" "1 1" "1 0"; do
  set -- $v
  NQA_STAGE1_PAIR=$1 NQA_STAGE1_A2_LINEAR=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r35_bench_p$1_l$2.log 2>&1
  echo "pair=$1 linear=$2"
  grep -o "\"value\": [0-9.]*\|\"stage1\": {[^}]*}\|sm_mhz\": [0-9.]*" gpurun_out/r35_bench_p$1_l$2.log
done
