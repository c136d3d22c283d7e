#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python bench.py --algo bfs --scale 24 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1 NT1024 ms/step %.3f'%(d['ms_per_step']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2964$N bench.py --gpus $N --algo bfs --scale 24 --steps 10 --warmup 3 \
      > gpurun_out/mg${N}_v3.json 2> gpurun_out/mg${N}_v3.err
python -c "import json,sys; d=json.load(open('gpurun_out/mg${N}_v3.json')); print('N=$N NT1024', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d['parity_vs_cpu_reference'])" || tail -8 gpurun_out/mg${N}_v3.err
