#!/bin/bash
# Round 2, session 3: A-DISTS parity subset + A-DISTS bench + window-kernel profile.   gpurun -- 'bash scripts/r3_run1.sh TAG'
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3b}
timeout 900 python -m pytest tests -x -q -m gpu -k "adists or small or shared_trunk or from_feats" > gpurun_out/${TAG}_t_adists.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/${TAG}_t_adists.log
timeout 600 python bench.py --workload adists --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_adists.json 2> gpurun_out/${TAG}_bench_adists.err; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/${TAG}_bench_adists.json').read().splitlines() if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(round(d['value'],1), 'e2e', (round(d['e2e']['value'],1) if d.get('e2e') else None), {k:round(v['ms_per_step'],1) for k,v in d['kernel_families'].items()})
PY
if [ "$2" = "prof" ]; then
CMD="python scripts/adists_once.py"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_mma" -c 2 -f -o gpurun_out/${TAG}_window $CMD > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_window.ncu-rep --page raw --csv > gpurun_out/${TAG}_window_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_window.ncu-rep --page source --csv > gpurun_out/${TAG}_window_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_window.ncu-rep
fi
