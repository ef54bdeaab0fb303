#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for v in 128 64; do
  NQA_PAIR_MINCIN=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r46_bench_mincin$v.log 2>&1
  echo "pair mincin=$v"
  grep -o "\"value\": [0-9.]*\|\"stage1\": {[^}]*}\|\"conv_tc\": {[^}]*}\|sm_mhz\": [0-9.]*" gpurun_out/r46_bench_mincin$v.log
done
NQA_PAIR_MINCIN=64 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or fused_stage1" 2>&1 | tail -2
