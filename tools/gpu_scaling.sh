#!/bin/bash
# torchrun bench at N GPUs (BFS and PageRank, RMAT-24).  Usage: tools/gpu_scaling.sh N [extra bench flags]
N=$1; shift
OUT=gpurun_out/r01
mkdir -p $OUT
for a in bfs pr; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2951$N bench.py --gpus $N --algo $a --scale 24 --steps 10 --warmup 3 "$@" \
      > $OUT/scaling_${a}_rmat24_g$N.json 2> $OUT/scaling_${a}_rmat24_g$N.err
  python -c "import json,sys; d=json.load(open('$OUT/scaling_${a}_rmat24_g$N.json')); print('$a', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']), d.get('parity_vs_cpu_reference'), d.get('max_rel_err'), d['roofline']['frac'])" || tail -5 $OUT/scaling_${a}_rmat24_g$N.err
done
