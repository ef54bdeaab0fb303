mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu_all22.log 2>&1; echo "all rc=$?"
tail -6 gpurun_out/t_gpu_all22.log
