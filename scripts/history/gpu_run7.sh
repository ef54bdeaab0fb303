mkdir -p gpurun_out
set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/bench_r1j_base.json 2>&1
NQA_CONV_HALO_MAXCIN=512 timeout 300 $B > gpurun_out/bench_r1j_h512.json 2>&1
NQA_CONV_HALO_MAXCIN=512 NQA_MICRO_IMAGES=16 timeout 300 $B > gpurun_out/bench_r1j_h512_m16.json 2>&1
NQA_MICRO_IMAGES=16 timeout 300 $B > gpurun_out/bench_r1j_m16.json 2>&1
NQA_PAIR_MINCIN=64 timeout 300 $B > gpurun_out/bench_r1j_p64.json 2>&1
NQA_CONV_HALO_MAXCIN=512 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1j_h512.csv python bench.py --steps 1 --warmup 3 --pairs 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r1j.log 2>&1; echo "ncu rc=$?"
grep -h '"value"' gpurun_out/bench_r1j_*.json | cut -c1-120
