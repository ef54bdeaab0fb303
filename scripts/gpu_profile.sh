#!/bin/bash
# Profile artifacts of the bench command (committed under profiles/ by hand afterwards):
#   launch list (gpu__time_duration per kernel) and an `ncu --set full` capture of one 8-pair micro-batch.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r1x}
timeout 300 python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_pairs8.log 2>&1 || { tail -3 gpurun_out/${TAG}_bench_pairs8.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_pairs8.csv python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"stage1_fused|pool_stats|pair_kernel|halo_kernel|stats_nhwc|stats_planar|dists_terms" -s 22 -c 24 -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep gpurun_out/${TAG}_ncu_full_summary.json
# gpurun copies back at most 64 MiB: keep the summary, drop the report
rm -f gpurun_out/${TAG}_full.ncu-rep
