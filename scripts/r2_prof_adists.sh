#!/bin/bash
# ncu artifacts of the A-DISTS step: launch list of one 4-pair micro-batch and a --set full capture of the window kernel.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2c}
CMD="python bench.py --workload adists --pairs 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/${TAG}_adists_p4.json 2>&1 || { tail -3 gpurun_out/${TAG}_adists_p4.json; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches_adists_pairs4.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_mma" -s 15 -c 5 -f -o gpurun_out/${TAG}_window_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${TAG}_window_full.ncu-rep
