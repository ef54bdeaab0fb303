#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q -k "pool or golden or micro_batch or flat or edge or 1080p or stats" > gpurun_out/r51_test.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r51_test.log
for v in 1 0; do
NQA_POOL_STATS_TMA=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r51_bench_tma$v.log 2>&1
echo "tma=$v"; grep -o "\"value\": [0-9.]*\|\"pool_stats\": {[^}]*}\|\"conv_tc\": {[^}]*}\|sm_mhz\": [0-9.]*" gpurun_out/r51_bench_tma$v.log
done
