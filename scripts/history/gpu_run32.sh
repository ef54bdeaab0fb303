mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1x_train.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1x_train.log 2>&1; echo "ncu rc=$?"
