mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/t_gpu_all19.log 2>&1; echo "all rc=$?"
grep -h "edge\|A-DISTS (\|DISTS (\|passed\|failed\|Error" gpurun_out/t_gpu_all19.log | tail -45
