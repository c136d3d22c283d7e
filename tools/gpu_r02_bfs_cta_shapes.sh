#!/bin/bash
mkdir -p gpurun_out
run() {  # name lib env
  GB200_LIB=$PWD/build/variants/$2 $3 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/bfsv_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1: bfs ms/step %.4f' % d['ms_per_step'])"
  GB200_BFS_TRACE=1 GB200_LIB=$PWD/build/variants/$2 $3 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "bfs trace" | tail -1
}
run nt1024x2 nt1024x2.so env
run nt768x2 nt768x2.so env
run nt512x3 nt512x3.so env
run nt1024x1 nt1024x1.so "env GB200_BFS_MINB=1"
