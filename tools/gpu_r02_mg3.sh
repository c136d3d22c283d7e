#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -m gpu 2>&1 | tail -3
GB200_DIST_ROW_WEIGHT=100000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2963$N bench.py --gpus $N --algo bfs --scale 24 --steps 10 --warmup 3 \
      > gpurun_out/mg${N}_v2.json 2> gpurun_out/mg${N}_v2.err
python -c "import json,sys; d=json.load(open('gpurun_out/mg${N}_v2.json')); print('N=$N', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['parity_vs_cpu_reference'])" || tail -8 gpurun_out/mg${N}_v2.err
