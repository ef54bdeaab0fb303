#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:stage1_fused -s 1 -c 1 python scripts/s1_once.py 2>&1 | grep -E "stage1_fused|gpu__time|pipe_tensor" 
NQA_STAGE1_PAIR=1 timeout 500 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:stage1_fused -s 1 -c 1 python scripts/s1_once.py 2>&1 | grep -E "stage1_fused|gpu__time|pipe_tensor"
