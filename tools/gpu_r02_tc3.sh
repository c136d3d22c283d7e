#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spgemmHashKernel -s 1 -c 1 -f \
    -o gpurun_out/prof_tchash_m python bench.py --algo tc --scale 22 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_ncu.py full gpurun_out/prof_tchash_m.ncu-rep 2>&1 | head -40
