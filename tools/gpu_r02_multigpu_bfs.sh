#!/bin/bash
# distributed fused BFS: shipped library vs a build with another CTA shape (GB200_LIB)
N=${1:-2}
mkdir -p gpurun_out
for lib in shipped build/variants/dist768.so; do
  if [ "$lib" = "shipped" ]; then unset GB200_LIB; else export GB200_LIB=$PWD/$lib; fi
  tag=$(basename $lib .so)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 2992$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg${N}_$tag.json 2> gpurun_out/mg${N}_$tag.err
  python -c "import json; d=json.load(open('gpurun_out/mg${N}_$tag.json')); print('N=$N $tag ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', d.get('parity_vs_cpu_reference'))" || tail -12 gpurun_out/mg${N}_$tag.err
done
