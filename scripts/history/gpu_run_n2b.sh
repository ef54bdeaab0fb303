mkdir -p gpurun_out
set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/video_nccl_check.py > gpurun_out/video_nccl_n2.log 2>&1; echo "video n2 rc=$?"
grep -h "world\|Error\|error" gpurun_out/video_nccl_n2.log | tail -5
timeout 600 python -m pytest tests/test_frmodel.py -q -m gpu > gpurun_out/t_frmodel.log 2>&1; echo "frmodel rc=$?"; tail -3 gpurun_out/t_frmodel.log
