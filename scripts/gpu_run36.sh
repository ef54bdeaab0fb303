#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 240 python scripts/s1_ablate.py > gpurun_out/r36_ablate.log 2>&1
cat gpurun_out/r36_ablate.log | tail -20
