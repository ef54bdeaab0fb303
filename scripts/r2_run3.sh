#!/bin/bash
# Round 2, run 3: whole GPU suite, smoke, the headline bench line, the video (configs[3]) and latency workloads.
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2m}
timeout 1500 python -m pytest tests -x -q -m gpu -rP > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
grep "\[parity\] resize" gpurun_out/${TAG}_tests.log | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --workload video > gpurun_out/${TAG}_bench_video_n1.json 2> gpurun_out/${TAG}_bench_video_n1.err; echo "video rc=$?"; tail -2 gpurun_out/${TAG}_bench_video_n1.err
timeout 300 python bench.py --workload latency > gpurun_out/${TAG}_bench_latency.json 2> gpurun_out/${TAG}_bench_latency.err; echo "latency rc=$?"; tail -2 gpurun_out/${TAG}_bench_latency.err
python - <<PY
import json
def last(f):
    l=[x for x in open(f).read().splitlines() if x.startswith('{')]
    return json.loads(l[-1]) if l else None
d=last('gpurun_out/${TAG}_bench_n1.json')
if d:
    print('dists', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
    a=d['adists']; print('adists', round(a['value'],1), 'e2e', round(a['e2e']['value'],1), {k:round(v['ms_per_step'],1) for k,v in a['kernel_families'].items()})
    t=d['train']; print('train', round(t['value'],1), t['roofline']['family'], t['roofline']['frac'])
v=last('gpurun_out/${TAG}_bench_video_n1.json')
if v: print('video', round(v['value'],1), 'ms', round(v['ms_per_step']), v['gathered_matches_rescored_sample'], v['config']['frames'], v['dists_vs_adists_over_videos'])
l=last('gpurun_out/${TAG}_bench_latency.json')
if l: print({k:(round(x['us_per_call_synchronised']),round(x['us_per_call_back_to_back'])) for k,x in l['calls'].items()})
PY
