#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 120 python scripts/adists_once.py > gpurun_out/r42_once.log 2>&1 || { tail -5 gpurun_out/r42_once.log; exit 1; }
# second call's window kernels: level 1 (64 ch, 1080p) is the 2nd window launch of a call (level 0 first)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_kernel -s 6 -c 2 -f -o gpurun_out/r42_window python scripts/adists_once.py > gpurun_out/r42_ncu.log 2>&1
tail -3 gpurun_out/r42_ncu.log
