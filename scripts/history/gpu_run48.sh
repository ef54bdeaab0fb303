#!/bin/bash
cd /root/repo
for v in "0 1" "0 0" "1 1" "1 0"; do
  set -- $v
  echo "pair=$1 linear=$2"
  NQA_STAGE1_PAIR=$1 NQA_STAGE1_A2_LINEAR=$2 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:stage1_fused -s 1 -c 1 python scripts/s1_once.py 2>&1 | grep -E "gpu__time|pipe_tensor"
done
