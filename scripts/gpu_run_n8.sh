#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
start=$(date +%s)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r1y_bench_n$N.json 2> gpurun_out/r1y_bench_n$N.err; echo "n$N rc=$? wall $(( $(date +%s) - start )) s"
grep -o "\"value\": [0-9.]*" gpurun_out/r1y_bench_n$N.json | head -3
tail -3 gpurun_out/r1y_bench_n$N.err
