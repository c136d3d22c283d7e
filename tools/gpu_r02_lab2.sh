#!/bin/bash
mkdir -p gpurun_out
L=build/lab/spmv_lab
echo "== sanitizer scale 13 v5 ==" > gpurun_out/lab6.log
timeout 300 compute-sanitizer --tool memcheck $L 13 16 1 2>&1 | grep -v "GB/s" >> gpurun_out/lab6.log
echo "== scale 22 ==" >> gpurun_out/lab6.log
timeout 300 $L 22 16 5 >> gpurun_out/lab6.log 2>&1
echo "rc=$?" >> gpurun_out/lab6.log
cat gpurun_out/lab6.log
