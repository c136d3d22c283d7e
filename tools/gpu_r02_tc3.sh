#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spgemmHashKernel -c 2 -f \
    -o gpurun_out/prof_tchash python bench.py --algo tc --scale 20 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/
python tools/summarize_ncu.py full gpurun_out/prof_tchash.ncu-rep 2>&1 | head -120
