mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_kernel" -s 1 -c 1 -o gpurun_out/prof_r1s_window python bench.py --workload adists --steps 1 --warmup 3 --pairs 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_r1s.log 2>&1; echo "ncu full rc=$?"
