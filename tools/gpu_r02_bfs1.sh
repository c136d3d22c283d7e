#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "bfs" > gpurun_out/bfs1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/bfs1_pytest.log
tail -15 gpurun_out/bfs1_pytest.log
timeout 300 python bench.py --algo bfs --scale 24 --steps 10 --warmup 3 > gpurun_out/bfs1_bench.json 2> gpurun_out/bfs1_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bfs1_bench.json"))
    print("bfs 24 ms/step %.3f e2e %.3f parity %s launches %d"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_vs_cpu_reference"], d["gpu_launches"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bfs1_bench.err").read()[-1500:])
PY
GB200_BFS_FUSED=0 timeout 300 python bench.py --algo bfs --scale 24 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('unfused ms/step %.3f launches %d'%(d['ms_per_step'], d['gpu_launches']))"
