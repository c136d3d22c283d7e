#!/bin/bash
# multi-GPU: dist tests + BFS bench at N GPUs (fused vs host-loop exchange)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -m gpu > gpurun_out/mg${N}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/mg${N}_pytest.log
tail -6 gpurun_out/mg${N}_pytest.log
for fused in 1 0; do
  GB200_DIST_BFS_FUSED=$fused timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2961$N bench.py --gpus $N --algo bfs --scale 24 --steps 10 --warmup 3 \
      > gpurun_out/mg${N}_bfs_fused$fused.json 2> gpurun_out/mg${N}_bfs_fused$fused.err
  python -c "import json,sys; d=json.load(open('gpurun_out/mg${N}_bfs_fused$fused.json')); print('N=$N fused=$fused', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity_vs_cpu_reference'), 'launches', d['gpu_launches'])" || tail -8 gpurun_out/mg${N}_bfs_fused$fused.err
done
