mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all11.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/t_gpu_all11.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_r1m.json
timeout 600 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_r1m_train.json 2> gpurun_out/bench_r1m_train.err; echo "train rc=$?"
cat gpurun_out/bench_r1m_train.json; tail -3 gpurun_out/bench_r1m_train.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1m_ref.json 2> gpurun_out/bench_r1m_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_r1m_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1m.csv python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1m.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stage1_fused|pool_stats|pair_kernel|halo_kernel|stats_nhwc|stats_planar" -c 22 -o gpurun_out/prof_r1m python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_r1m.log 2>&1; echo "ncu full rc=$?"
