#!/bin/bash
# multi-GPU check after the pull changes of the distributed kernel: tests, then bench
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2972$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg${N}_v7.json 2> gpurun_out/mg${N}_v7.err
python -c "import json; d=json.load(open('gpurun_out/mg${N}_v7.json')); print('N=$N ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity_vs_cpu_reference'))" || tail -12 gpurun_out/mg${N}_v7.err
GB200_BFS_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2973$N bench.py --gpus $N --steps 2 --warmup 2 --no-cpu-baseline > /dev/null 2> gpurun_out/mg${N}_v7t.err
grep "rank 0 level" gpurun_out/mg${N}_v7t.err | tail -6
for a in pr sssp; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2974$N bench.py --gpus $N --algo $a --scale 24 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/mg${N}_$a.json 2> gpurun_out/mg${N}_$a.err
python -c "import json; d=json.load(open('gpurun_out/mg${N}_$a.json')); print('N=$N $a rmat24 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))" || tail -12 gpurun_out/mg${N}_$a.err
done
