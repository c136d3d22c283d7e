#!/bin/bash
mkdir -p gpurun_out
for mb in 4 3 2; do
GB200_BFS_MINB=$mb timeout 300 python bench.py --algo bfs --scale 24 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('minb $mb ms/step %.3f launches %d'%(d['ms_per_step'], d['gpu_launches']))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bfsFusedKernel -s 3 -c 1 -o gpurun_out/bfs_prof -f python bench.py --algo bfs --scale 24 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bfs_ncu.log 2>&1
tail -3 gpurun_out/bfs_ncu.log
