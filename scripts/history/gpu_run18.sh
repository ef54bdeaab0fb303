mkdir -p gpurun_out
timeout 300 python scripts/dbg_adists_small.py > gpurun_out/dbg_adists_small.log 2>&1; cat gpurun_out/dbg_adists_small.log | tail -12
