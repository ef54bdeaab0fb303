#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma_fed" 2>&1 | tail -3
timeout 300 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -o "\"value\": [0-9.]*\|\"pool_stats\": {[^}]*}"
NQA_POOL_STATS_TMA=0 timeout 300 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -o "\"value\": [0-9.]*\|\"pool_stats\": {[^}]*}"
