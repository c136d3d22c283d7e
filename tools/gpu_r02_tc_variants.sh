#!/bin/bash
# kernel-parameter variants of the hash SpGEMM (build/variants/*.so), RMAT-22
# (profiles/r02_tc_hash_variants.txt).  Built on the CPU box with the library's nvcc
# line plus, e.g.:
#   v1: -DGB_HASH_SLOTS_M=2048 -DGB_HASH_CTAS_M=7 -DGB_HASH_SLOTS_L=8192 -DGB_HASH_SEG_L=4096 -DGB_HASH_CTAS_L=2 -DGB_HASH_UNROLL_L=8
#   v2: -DGB_HASH_SLOTS_M=2048 -DGB_HASH_CTAS_M=7 -DGB_HASH_UNROLL_L=8            (shipped defaults)
#   v3: v1 with -DGB_HASH_SEG_L=6144 -DGB_HASH_UNROLL_L=4 -DGB_HASH_CHUNK_L=1024 -DGB_HASH_UNROLL_M=2
# and selected at run time through GB200_LIB.
mkdir -p gpurun_out
for v in $(ls build/variants/*.so); do
  name=$(basename $v .so)
  GB200_LIB=$PWD/$v timeout 600 python bench.py --algo tc --scale 22 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/tcv_$name.json 2> gpurun_out/tcv_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/tcv_$name.json"))
    print("$name: ms %.3f tris %s" % (d["ms_per_step"], d.get("triangles", d.get("result"))))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/tcv_$name.err").read()[-800:])
PY
  GB200_LIB=$PWD/$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/tcv_$name.csv python bench.py --algo tc --scale 22 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/tcv_$name.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
out=[float(r[vi])/1e6 for r in rows[1:] if 'spgemmHashKernel' in r[ki]][:6]
print("   L,M,S pass1 | L,M,S pass2 (ms):", " ".join("%.2f"%x for x in out))
PY
done
