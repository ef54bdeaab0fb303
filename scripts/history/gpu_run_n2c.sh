#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
start=$(date +%s)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1x_bench_n2.json 2> gpurun_out/r1x_bench_n2.err; echo "n2 rc=$? wall $(( $(date +%s) - start )) s"
grep -o "\"value\": [0-9.]*" gpurun_out/r1x_bench_n2.json | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/r1x_bench_n2_ref.json 2> gpurun_out/r1x_bench_n2_ref.err; echo "n2 ref rc=$?"
cut -c1-300 gpurun_out/r1x_bench_n2_ref.json
