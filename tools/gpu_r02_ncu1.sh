#!/bin/bash
mkdir -p gpurun_out
L=build/lab/spmv_lab
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmvHubKernel -s 2 -c 1 -o gpurun_out/hub_prof -f $L 22 16 1 > gpurun_out/ncu1.log 2>&1
echo "rc=$?" >> gpurun_out/ncu1.log
tail -5 gpurun_out/ncu1.log
ls -la gpurun_out/
