#!/bin/bash
# Re-run of the bench lines whose roofline block changed with bench.py (burst peak for a minority conv family), final code.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3z}
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "$name rc=$?"; }
run adists --workload adists --steps 5 --warmup 3
run both --workload both --steps 5 --warmup 3 --no-cpu-baseline
python - <<PY
import json,glob
for f in ('n1','adists','both'):
    l=[x for x in open('gpurun_out/${TAG}_bench_%s.json'%f).read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); r=d.get('roofline') or {}
    print(f.ljust(8), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'top', r.get('family'), 'frac', round(r.get('frac'),3), 'traffic', r.get('traffic'), [ (o['family'], round(o['frac'],3)) for o in d.get('roofline_other',[])])
    for k in ('adists','train'):
        if isinstance(d.get(k), dict): print('   ', k, round(d[k]['value'],1), (d[k].get('e2e') or {}).get('value'), (d[k].get('roofline') or {}).get('family'), (d[k].get('roofline') or {}).get('frac'))
PY
