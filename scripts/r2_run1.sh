#!/bin/bash
# Round 2, run 1: full GPU suite with the new A-DISTS / stream tests, the headline bench, an A-DISTS launch list.
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2a_tests.log
grep "\[parity\] A-DISTS 1080p" gpurun_out/r2a_tests.log | head
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2a_bench_n1.err
timeout 300 python bench.py --workload adists --pairs 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_adists_p4.json 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches_adists_pairs4.csv python bench.py --workload adists --pairs 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_ncu_list.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2a_bench_n1.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1])
print('dists', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
a=d['adists']; print('adists', round(a['value'],1), 'e2e', a['e2e'], {k:round(v['ms_per_step'],1) for k,v in a['kernel_families'].items()})
print('adists roofline', json.dumps(a['roofline'])[:900])
t=d['train']; print('train', round(t['value'],1), t['roofline']['frac'])
PY
