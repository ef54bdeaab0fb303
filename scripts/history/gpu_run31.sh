mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/t_31.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_31.log
