#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for rw in 0 32 128 100000; do
  GB200_DIST_ROW_WEIGHT=$rw timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2962$N bench.py --gpus $N --algo bfs --scale 24 --steps 10 --warmup 3 --no-cpu-baseline \
      > gpurun_out/mg${N}_rw$rw.json 2> gpurun_out/mg${N}_rw$rw.err
  python -c "import json,sys; d=json.load(open('gpurun_out/mg${N}_rw$rw.json')); print('N=$N row_weight=$rw', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['config']['partition'][-60:])" || tail -8 gpurun_out/mg${N}_rw$rw.err
done
