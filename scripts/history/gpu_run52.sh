#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_stats -s 4 -c 1 -f -o gpurun_out/r52_pool_tma python scripts/s1_once.py > gpurun_out/r52_ncu.log 2>&1
NQA_POOL_STATS_TMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_stats -s 4 -c 1 -f -o gpurun_out/r52_pool_reg python scripts/s1_once.py > gpurun_out/r52_ncu2.log 2>&1
ls gpurun_out/r52*
