#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_stage1" 2>&1 | tail -2
timeout 500 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:stage1_fused -s 1 -c 1 python scripts/s1_once.py 2>&1 | grep -E "gpu__time|pipe_tensor"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r47_bench.log 2>&1
grep -o "\"value\": [0-9.]*\|\"stage1\": {[^}]*}\|\"conv_tc\": {[^}]*}\|sm_mhz\": [0-9.]*" gpurun_out/r47_bench.log
