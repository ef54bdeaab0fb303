#!/bin/bash
# Profile artifacts of the final code (copied into profiles/ afterwards): launch lists of one DISTS and one A-DISTS
# micro-batch and `ncu --set full` captures of their kernels.   gpurun -- 'bash scripts/r2_profile.sh TAG'
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2z}
D="python bench.py --workload dists --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e"
A="python bench.py --workload adists --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e"
timeout 300 $D > gpurun_out/${TAG}_dists_p8.log 2>&1 || { tail -3 gpurun_out/${TAG}_dists_p8.log; exit 1; }
timeout 300 $A > gpurun_out/${TAG}_adists_p8.log 2>&1 || { tail -3 gpurun_out/${TAG}_adists_p8.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_dists_pairs8.csv $D > /dev/null 2>&1; echo "list dists rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_adists_pairs8.csv $A > /dev/null 2>&1; echo "list adists rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"stage1_fused|pool_stats|pair_kernel|halo_kernel|stats_nhwc|stats_planar|dists_terms" -s 22 -c 24 -f -o gpurun_out/${TAG}_dists_full $D > /dev/null 2>&1; echo "full dists rc=$?"
python scripts/ncu_summary.py gpurun_out/${TAG}_dists_full.ncu-rep gpurun_out/${TAG}_dists_ncu_full_summary.json | cut -c1-170
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_|planes_sums|l2pool_nhwc|gamma_stats|prod_|reduce_splits|chan_kernel" -s 40 -c 30 -f -o gpurun_out/${TAG}_adists_full $A > /dev/null 2>&1; echo "full adists rc=$?"
python scripts/ncu_summary.py gpurun_out/${TAG}_adists_full.ncu-rep gpurun_out/${TAG}_adists_ncu_full_summary.json | cut -c1-170
rm -f gpurun_out/${TAG}_dists_full.ncu-rep
ls -la gpurun_out/${TAG}_adists_full.ncu-rep
