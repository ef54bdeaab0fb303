mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "large_frames or invalid_inputs" > gpurun_out/t_23.log 2>&1; echo "rc=$?"
tail -15 gpurun_out/t_23.log
