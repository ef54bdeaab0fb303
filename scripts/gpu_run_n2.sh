mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_r1k_n2.json 2> gpurun_out/bench_r1k_n2.err; echo "n2 rc=$?"
cat gpurun_out/bench_r1k_n2.json; tail -5 gpurun_out/bench_r1k_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 -m pytest tests/test_video_gloo.py -q -x > gpurun_out/t_video_n2.log 2>&1; tail -3 gpurun_out/t_video_n2.log
