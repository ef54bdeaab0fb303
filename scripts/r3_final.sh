#!/bin/bash
# Final single-GPU run of the round: GPU suite, smoke, the driver's default bench line, the reference arm and every other
# workload.   gpurun -- 'bash scripts/r3_final.sh TAG'
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3z}
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/${TAG}_t_gpu.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${TAG}_t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "reference arm rc=$?"
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err; echo "$name rc=$?"; }
run adists --workload adists --steps 5 --warmup 3
run both --workload both --steps 5 --warmup 3 --no-cpu-baseline
run video_n1 --workload video --no-cpu-baseline
run latency --workload latency
run train --workload train --steps 5 --warmup 3 --no-cpu-baseline
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*.json')):
    l=[x for x in open(f).read().splitlines() if x.startswith('{')]
    if not l: print(f,'NO LINE'); continue
    d=json.loads(l[-1]); r=d.get('roofline') or {}
    print(f.split('bench_')[1][:-5].ljust(14), d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'top', r.get('family'), 'frac', r.get('frac'), 'traffic', r.get('traffic'))
    for k in ('adists','train'):
        if isinstance(d.get(k), dict): print('   ', k, d[k].get('value'), (d[k].get('e2e') or {}).get('value'))
PY
