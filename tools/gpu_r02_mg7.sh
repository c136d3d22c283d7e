#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -m gpu 2>&1 | tail -2; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2968$N bench.py --gpus $N --algo bfs --scale 24 --steps 20 --warmup 5 \
      > gpurun_out/mg${N}_v6.json 2> gpurun_out/mg${N}_v6.err
python -c "import json,sys; d=json.load(open('gpurun_out/mg${N}_v6.json')); print('N=$N', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['parity_vs_cpu_reference'])" || tail -8 gpurun_out/mg${N}_v6.err
GB200_BFS_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2969$N bench.py --gpus $N --algo bfs --scale 24 --steps 3 --warmup 3 --no-cpu-baseline \
      > /dev/null 2> gpurun_out/mg${N}_v6t.err
grep "^rank 0" gpurun_out/mg${N}_v6t.err | tail -6
