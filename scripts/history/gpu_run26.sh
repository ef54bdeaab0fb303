mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1r_adists.csv python bench.py --workload adists --steps 1 --warmup 3 --pairs 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1r_adists.log 2>&1; echo "ncu rc=$?"
