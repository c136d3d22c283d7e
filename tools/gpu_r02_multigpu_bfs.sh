#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2982$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg${N}_v7.json 2> gpurun_out/mg${N}_v7.err
python -c "import json; d=json.load(open('gpurun_out/mg${N}_v7.json')); print('N=$N ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity_vs_cpu_reference'))" || tail -12 gpurun_out/mg${N}_v7.err
