#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --algo bfs --scale 24 --steps 10 --warmup 3 > gpurun_out/t3_bfs.json 2> gpurun_out/t3_bfs.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t3_bfs.json"))
    r=d["roofline"]
    print("bfs 24 ms/step %.3f e2e %.3f parity %s launches %d frac %.3f share %.2f"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_vs_cpu_reference"], d["gpu_launches"], r["frac"], r["share_of_step"]), r.get("fused_traversal"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/t3_bfs.err").read()[-1500:])
PY
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/t3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t3_pytest.log
tail -25 gpurun_out/t3_pytest.log
