mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/t_gpu_all12.log 2>&1; echo "all rc=$?"
grep -h "flavour\|passed\|failed\|Error" gpurun_out/t_gpu_all12.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke12.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke12.log
