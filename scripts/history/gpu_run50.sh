#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r50_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r50_tests.log
for v in 1 0; do
NQA_STAGE1_PAIR=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r50_bench_p$v.log 2>&1
echo "pair=$v"; grep -o "\"value\": [0-9.]*\|\"stage1\": {[^}]*}\|\"conv_tc\": {[^}]*}\|sm_mhz\": [0-9.]*" gpurun_out/r50_bench_p$v.log
done
