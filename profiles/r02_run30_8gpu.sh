#!/bin/bash
# round 2, final code: the 48-scene job on 8 / 4 / 2 GPUs of one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P=29631
for N in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
P=$((P+1))
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value'],3), round(d['ms_per_step'],1), d['balance'], round(d['e2e']['value'],3))
"
done
