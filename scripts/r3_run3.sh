#!/bin/bash
# Round 2, session 3: the strip-plane A-DISTS path, mode by mode (NQA_ADISTS_STRIPS = 1: strips from the NHWC level, 2: from the
# convolution epilogues), each with the A-DISTS tests and the bench.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3d}
for mode in 1 2; do
  NQA_ADISTS_STRIPS=$mode timeout 900 python -m pytest tests -x -q -m gpu -k "adists or small or shared_trunk" > gpurun_out/${TAG}_t_adists_m$mode.log 2>&1; echo "mode $mode tests rc=$?"
  tail -4 gpurun_out/${TAG}_t_adists_m$mode.log
done
for mode in 0 2; do
NQA_ADISTS_STRIPS=$mode timeout 600 python bench.py --workload adists --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_adists_m$mode.json 2> gpurun_out/${TAG}_bench_adists_m$mode.err; echo "bench mode $mode rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench_adists_m$mode.json').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(round(d['value'],1), 'e2e', (round(d['e2e']['value'],1) if d.get('e2e') else None), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
done
