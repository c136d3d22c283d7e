#!/bin/bash
# hash-formulation masked SpGEMM: parity tests, then timings against the search kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "tc or triangle or mxm" 2>&1 | tail -5
for sc in 18 20 22; do
  for h in 1 0; do
    GB200_SPGEMM_HASH=$h timeout 600 python bench.py --algo tc --scale $sc --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/tc_$sc_$h.json 2> gpurun_out/tc_$sc_$h.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/tc_$sc_$h.json"))
    print("scale $sc hash=$h: ms %.3f parity %s launches/step %.1f per_mxv %s" % (d["ms_per_step"], d["parity_vs_cpu_reference"], d["gpu_launches"]/d["steps"], json.dumps(d.get("per_mxv"))[:600]))
except Exception as e:
    print("scale $sc hash=$h failed", e); print(open("gpurun_out/tc_$sc_$h.err").read()[-1500:])
PY
  done
done
