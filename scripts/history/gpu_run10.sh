mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gradient or training_step" -s > gpurun_out/t_bwd.log 2>&1; echo "bwd rc=$?"
grep -h "parity\|passed\|failed\|Error\|error" gpurun_out/t_bwd.log | tail -30
