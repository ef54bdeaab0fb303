#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
export NQA_STAGE1_PAIR=1
timeout 120 python scripts/s1_once.py > gpurun_out/r43_once.log 2>&1 || { tail -5 gpurun_out/r43_once.log; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:stage1_fused -s 1 -c 1 -f -o gpurun_out/r43_stage1_pair python scripts/s1_once.py > gpurun_out/r43_ncu.log 2>&1
tail -3 gpurun_out/r43_ncu.log
