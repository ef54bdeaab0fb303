#!/bin/bash
# Round 2, session 3: full GPU suite, smoke, the default bench line (what the driver runs) and the reference arm.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3c}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_t_gpu.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/${TAG}_t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench_n1.json').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(round(d['value'],1), 'e2e', (round(d['e2e']['value'],1) if d.get('e2e') else None), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
    for k in ('adists','train'):
        if d.get(k): print(k, round(d[k]['value'],1), 'e2e', (d[k].get('e2e') or {}).get('value'))
    print('cpu', d.get('cpu_baseline'))
PY
