#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host" 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r57_bench.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r57_bench.log').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['value']/d['value'],4))
PY
