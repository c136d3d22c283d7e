#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2966$N bench.py --gpus $N --algo bfs --scale 24 --steps 10 --warmup 3 \
      > gpurun_out/mg${N}_v5.json 2> gpurun_out/mg${N}_v5.err
python -c "import json,sys; d=json.load(open('gpurun_out/mg${N}_v5.json')); print('N=$N', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['parity_vs_cpu_reference'])" || tail -8 gpurun_out/mg${N}_v5.err
GB200_BFS_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2967$N bench.py --gpus $N --algo bfs --scale 24 --steps 3 --warmup 3 --no-cpu-baseline \
      > /dev/null 2> gpurun_out/mg${N}_v5t.err
grep "^rank" gpurun_out/mg${N}_v5t.err | tail -12
