mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
cat gpurun_out/bench_final.json
timeout 600 python bench.py --workload adists --steps 3 --warmup 3 > gpurun_out/bench_final_adists.json 2> gpurun_out/bench_final_adists.err; echo "adists rc=$?"
timeout 600 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_final_train.json 2> gpurun_out/bench_final_train.err; echo "train rc=$?"
python - <<'PY'
import json
for f in ('bench_final_ref','bench_final','bench_final_adists','bench_final_train'):
    l=[x for x in open(f'gpurun_out/{f}.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print(f, d['metric'], round(d['value'],2), 'e2e', (round(d['e2e']['value'],2) if d.get('e2e') else None), 'cpu', d.get('cpu_baseline'))
PY
