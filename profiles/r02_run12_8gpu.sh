#!/bin/bash
# round 2, GPU call 12 (8 B200): the 48-scene job sharded over 8 / 4 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
python -c "
import json
for f in ('r02_bench_8gpu','r02_bench_4gpu'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['n_gpus'], round(d['value'],3), round(d['ms_per_step'],1), d['balance'], d['e2e']['value'])
"
tail -3 gpurun_out/r02_bench_8gpu.err
