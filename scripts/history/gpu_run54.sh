#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma_fed or golden or 1080p" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pool_stats -s 4 -c 1 python scripts/s1_once.py 2>&1 | grep -E "gpu__time|issue_active"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r54_bench.log 2>&1
grep -o "\"value\": [0-9.]*\|\"pool_stats\": {[^}]*}\|\"conv_tc\": {[^}]*}\|\"stage1\": {[^}]*}\|sm_mhz\": [0-9.]*" gpurun_out/r54_bench.log
