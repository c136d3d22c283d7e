#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/t2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t2_pytest.log
tail -12 gpurun_out/t2_pytest.log
for cfg in "tc 22" "tc 18"; do
  set -- $cfg
  timeout 900 python bench.py --algo $1 --scale $2 --steps 3 --warmup 3 > gpurun_out/t2_bench_$1_$2.json 2> gpurun_out/t2_bench_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t2_bench_$1_$2.json"))
    r=d["roofline"]
    print("$1 $2", "ms/step %.3f"%d["ms_per_step"], "parity", d["parity_vs_cpu_reference"], d.get("triangles"), d.get("triangle_check"), d["cpu_baseline"])
except Exception as e:
    print("$1 $2 failed", e); print(open("gpurun_out/t2_bench_$1_$2.err").read()[-1500:])
PY
done
timeout 600 python bench.py --impl reference --algo tc --scale 22 --steps 2 --warmup 1 2>/dev/null | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400
