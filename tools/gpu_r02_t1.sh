#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/t1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t1_pytest.log
tail -15 gpurun_out/t1_pytest.log
for cfg in "sssp 22" "pr 22" "sssp 24" "bfs 24"; do
  set -- $cfg
  timeout 600 python bench.py --algo $1 --scale $2 --steps 5 --warmup 3 > gpurun_out/t1_bench_$1_$2.json 2> gpurun_out/t1_bench_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t1_bench_$1_$2.json"))
    r=d["roofline"]
    print("$1 $2", "ms/step %.3f"%d["ms_per_step"], "parity", d["parity_vs_cpu_reference"], "dom", r["kernel"][:20], "ms/launch %.3f frac %.3f share %.2f"%(r["ms_per_launch"], r["frac"], r["share_of_step"]), "launches", d["gpu_launches"])
except Exception as e:
    print("$1 $2 failed", e)
PY
done
