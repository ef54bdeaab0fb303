mkdir -p gpurun_out
set -x
python - <<'PY' > gpurun_out/h2d_bw.log 2>&1
import torch, time
x = torch.empty(1<<30, dtype=torch.uint8).pin_memory()
d = torch.empty(1<<30, dtype=torch.uint8, device='cuda')
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print(f"pinned H2D {1.0737/dt:.1f} GB/s")
PY
cat gpurun_out/h2d_bw.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dists_64 or odd_45x77 or fused_stage1 or (gradients_through and fp32 and shape1) or adists_96x128 or flavours" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/sanitizer_memcheck.log
