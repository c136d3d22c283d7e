#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "bfs or BFS or gbfs" 2>&1 | tail -3 | cut -c1-200
GB200_BFS_TRACE=1 timeout 600 python bench.py --steps 6 --warmup 3 2> gpurun_out/bfs_trace.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bfs ms/step %.3f e2e %.3f parity %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['parity_vs_cpu_reference']))"
grep "bfs trace" gpurun_out/bfs_trace.err | tail -2
