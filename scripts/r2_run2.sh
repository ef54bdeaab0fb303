#!/bin/bash
# Round 2, run 2: the tensor-pipe A-DISTS window kernel -- parity (level by level at 1080p), then the A-DISTS bench.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 900 python -m pytest tests/test_gpu_adists_parity.py tests/test_gpu_parity.py tests/test_gpu_ops.py -x -q -m gpu -k "adists or shared_trunk" -rP > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
grep "\[parity\]" gpurun_out/${TAG}_tests.log | grep -i "adists\|A-DISTS" | head -60
timeout 600 python bench.py --workload adists --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_adists.json 2> gpurun_out/${TAG}_bench_adists.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_adists.err
NQA_ADISTS_MMA=0 timeout 600 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_adists_nomma.json 2>&1
python - <<PY
import json
for f in ('${TAG}_bench_adists','${TAG}_bench_adists_nomma'):
    l=[x for x in open(f'gpurun_out/{f}.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print(f, round(d['value'],1), 'e2e', (round(d['e2e']['value'],1) if d.get('e2e') else None), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
    r=d['roofline']; print('  top', r['family'], 'frac', r['frac'], r.get('hbm',{}).get('frac'))
PY
NQA_MICRO_IMAGES=32 timeout 600 python bench.py --workload adists --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_adists_mi32.json 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench_adists_mi32.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print('micro_images=32', round(d['value'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
