#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "sssp or pr or pagerank or dropin or gpr or gsssp" 2>&1 | tail -4 | cut -c1-200
for a in sssp pr; do
  for f in 1 0; do
  GB200_LOOP_STEPS=$f timeout 600 python bench.py --algo $a --steps 10 --warmup 3 > gpurun_out/e2e_$a$f.json 2> gpurun_out/e2e_$a$f.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/e2e_$a$f.json"))
    print("$a loop_steps=$f: ms/step %.3f value %.0f | e2e ms %.3f | launches/step %.1f parity %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]/d["steps"], d["parity_vs_cpu_reference"]), d.get("pagerank_check"))
except Exception as e:
    print("$a failed", e); print(open("gpurun_out/e2e_$a$f.err").read()[-1200:])
PY
  done
done
