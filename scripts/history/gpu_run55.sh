#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q -k "adists or as_map or layout or planes or shared_trunk" 2>&1 | tail -2
timeout 300 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r55_bench_adists.log 2>&1
grep -o "\"value\": [0-9.]*\|\"layout\": {[^}]*}\|\"adists\": {[^}]*}" gpurun_out/r55_bench_adists.log
