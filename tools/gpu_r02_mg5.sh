#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
GB200_BFS_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2965$N bench.py --gpus $N --algo bfs --scale 24 --steps 3 --warmup 3 --no-cpu-baseline \
      > gpurun_out/mg${N}_v4.json 2> gpurun_out/mg${N}_v4.err
python -c "import json,sys; d=json.load(open('gpurun_out/mg${N}_v4.json')); print('N=$N', 'ms', round(d['ms_per_step'],3))"
grep "^rank" gpurun_out/mg${N}_v4.err | tail -$((N*7))
