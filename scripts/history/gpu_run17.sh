mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "edge_shapes or non_contiguous" -s > gpurun_out/t_edge17.log 2>&1; echo "edge rc=$?"
grep -h "parity\|passed\|failed\|Error\|error" gpurun_out/t_edge17.log | tail -50
