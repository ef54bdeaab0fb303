#!/bin/bash
# `ncu --set full` capture (with source) of the level-1 / level-2 tensor-pipe window kernel of one 2-pair 1080p A-DISTS call;
# the raw and source pages are exported on the box.   gpurun -- 'bash scripts/r2_prof_window.sh TAG'
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3a}
CMD="python scripts/adists_once.py"
timeout 300 $CMD > gpurun_out/${TAG}_once.log 2>&1 || { tail -3 gpurun_out/${TAG}_once.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_mma" -c 2 -f -o gpurun_out/${TAG}_window $CMD > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_window.ncu-rep --page raw --csv > gpurun_out/${TAG}_window_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_window.ncu-rep --page source --csv > gpurun_out/${TAG}_window_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_window*
