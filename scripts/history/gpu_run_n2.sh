mkdir -p gpurun_out
set -x
date +%s > gpurun_out/n2_t0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_r1t_n2.json 2> gpurun_out/bench_r1t_n2.err; echo "n2 rc=$?"
date +%s > gpurun_out/n2_t1
python -c "
import json
l=[x for x in open('gpurun_out/bench_r1t_n2.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print(d['n_gpus'], d['value'], d['e2e']['value'] if d['e2e'] else None, d['ms_per_step'])"
echo "elapsed $(( $(cat gpurun_out/n2_t1) - $(cat gpurun_out/n2_t0) )) s"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_r1t_ref_n2.json 2> gpurun_out/bench_r1t_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/bench_r1t_ref_n2.json
