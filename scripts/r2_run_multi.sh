#!/bin/bash
# bench.py under torchrun on N GPUs of one box: the video workload (BASELINE configs[3], strong scaling) and the
# headline line (weak scaling).   gpurun --gpus N -- 'bash scripts/r2_run_multi.sh N TAG'
cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-r2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $RUN bench.py --gpus $N --workload video > gpurun_out/${TAG}_bench_video_n${N}.json 2> gpurun_out/${TAG}_bench_video_n${N}.err; echo "video rc=$?"; tail -2 gpurun_out/${TAG}_bench_video_n${N}.err
timeout 900 $RUN bench.py --gpus $N --steps 5 --warmup 3 --headline-only > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
def last(f):
    l=[x for x in open(f).read().splitlines() if x.startswith('{')]
    return json.loads(l[-1]) if l else None
v=last('gpurun_out/${TAG}_bench_video_n${N}.json')
if v: print('video n=$N', round(v['value'],1), 'pairs/s, ms', round(v['ms_per_step']), 'check', v['gathered_matches_rescored_sample'], 'checksum', v['score_checksum'])
d=last('gpurun_out/${TAG}_bench_n${N}.json')
if d: print('dists n=$N', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))
PY
