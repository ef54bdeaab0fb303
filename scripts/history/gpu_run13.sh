mkdir -p gpurun_out
set -x
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1n_adists.csv python bench.py --workload adists --steps 1 --warmup 3 --pairs 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1n_adists.log 2>&1; echo "ncu rc=$?"
