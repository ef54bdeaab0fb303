#!/bin/bash
# Profile artifacts of the final code (copied into profiles/ afterwards): launch lists of one 8-pair 1080p DISTS and A-DISTS
# call (ncu gpu__time_duration pass) and `ncu --set full` summaries of all their kernels.
#   gpurun -- 'bash scripts/r3_profile.sh TAG'
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3z}
for wl in dists adists; do
  CMD="python scripts/profile_once.py $wl 8"
  timeout 300 $CMD > gpurun_out/${TAG}_${wl}_once.log 2>&1 || { tail -3 gpurun_out/${TAG}_${wl}_once.log; exit 1; }
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_${wl}_pairs8.csv $CMD > /dev/null 2>&1; echo "list $wl rc=$?"
  timeout 1200 ncu --profile-from-start off --set full --clock-control none -f -o gpurun_out/${TAG}_${wl}_full $CMD > /dev/null 2>&1; echo "full $wl rc=$?"
  python scripts/ncu_summary.py gpurun_out/${TAG}_${wl}_full.ncu-rep gpurun_out/${TAG}_${wl}_ncu_full_summary.json | cut -c1-150
  rm -f gpurun_out/${TAG}_${wl}_full.ncu-rep
done
