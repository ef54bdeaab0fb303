#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1z_launches_adists_pairs4.csv python bench.py --workload adists --pairs 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r56_ncu_adists.log 2>&1
echo rc=$?
