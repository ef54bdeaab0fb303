#!/bin/bash
# `ncu --set full` (with source) of kernels matching REGEX in one 2-pair 1080p A-DISTS call; raw + source pages exported.
#   gpurun -- 'bash scripts/r3_prof_kernel.sh TAG REGEX [COUNT]'
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3e}; RX=${2:-pool_strips}; N=${3:-2}
CMD=${4:-"python scripts/adists_once.py"}
timeout 300 $CMD > gpurun_out/${TAG}_once.log 2>&1 || { tail -3 gpurun_out/${TAG}_once.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -c $N -f -o gpurun_out/${TAG}_k $CMD > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_k.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_k.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_k.ncu-rep
