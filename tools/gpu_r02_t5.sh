#!/bin/bash
# full GPU suite + TC timing after the hash SpGEMM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | cut -c1-200
for sc in 20 22; do
  timeout 600 python bench.py --algo tc --scale $sc --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/tc5_$sc.json 2> gpurun_out/tc5_$sc.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/tc5_$sc.json"))
    print("tc scale $sc: ms %.3f parity %s launches/step %.1f" % (d["ms_per_step"], d["parity_vs_cpu_reference"], d["gpu_launches"]/d["steps"]))
except Exception as e:
    print("tc scale $sc failed", e); print(open("gpurun_out/tc5_$sc.err").read()[-800:])
PY
done
