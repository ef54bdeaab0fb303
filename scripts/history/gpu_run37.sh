#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 120 python scripts/s1_once.py > gpurun_out/r37_once.log 2>&1 || exit 1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:stage1_fused -s 1 -c 1 -f -o gpurun_out/r37_stage1 python scripts/s1_once.py > gpurun_out/r37_ncu.log 2>&1
tail -3 gpurun_out/r37_ncu.log
ls -la gpurun_out/
