#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_debug.py 2>&1 | tail -30
timeout 300 python -m pytest tests -q -m gpu -k "triangle or gtc or tc" 2>&1 | tail -40 | cut -c1-200
for sc in 20 22; do
  for h in 1 0; do
    GB200_SPGEMM_HASH=$h timeout 600 python bench.py --algo tc --scale $sc --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/tc_${sc}_$h.json 2> gpurun_out/tc_${sc}_$h.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/tc_${sc}_$h.json"))
    print("scale $sc hash=$h: ms %.3f parity %s" % (d["ms_per_step"], d["parity_vs_cpu_reference"]))
except Exception as e:
    print("scale $sc hash=$h failed", e); print(open("gpurun_out/tc_${sc}_$h.err").read()[-1500:])
PY
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/tc22_launches.csv python bench.py --algo tc --scale 22 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_ncu.py launches gpurun_out/tc22_launches.csv 2>&1 | head -12
