mkdir -p gpurun_out
timeout 300 python scripts/dbg_1x1.py > gpurun_out/dbg_1x1.log 2>&1; tail -25 gpurun_out/dbg_1x1.log
