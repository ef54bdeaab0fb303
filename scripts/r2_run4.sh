#!/bin/bash
# Round 2, run 4: gradient tests (bf16x3 backward), then the training bench in the three precisions.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2q}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -rP -k "gradient or backward or saved or tape" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
grep "\[parity\] d score/d x through the trunk" gpurun_out/${TAG}_tests.log | cut -c1-150
for P in bf16 bf16x3; do
timeout 600 python bench.py --workload train --precision $P --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_train_$P.json 2> gpurun_out/${TAG}_bench_train_$P.err; echo "train $P rc=$?"; tail -2 gpurun_out/${TAG}_bench_train_$P.err
done
python - <<PY
import json
for P in ('bf16','bf16x3'):
    l=[x for x in open(f'gpurun_out/${TAG}_bench_train_{P}.json').read().splitlines() if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print('train',P, round(d['value'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()}, d['roofline']['family'], d['roofline']['frac'])
PY
