#!/bin/bash
mkdir -p gpurun_out
L=build/lab/spmv_lab
echo "== sanitizer scale 13 v3 ==" > gpurun_out/lab3.log
timeout 300 compute-sanitizer --tool memcheck $L 13 16 1 2>&1 | grep -v "GB/s" >> gpurun_out/lab3.log
echo "== scale 22 ==" >> gpurun_out/lab3.log
timeout 300 $L 22 16 5 >> gpurun_out/lab3.log 2>&1
echo "rc=$?" >> gpurun_out/lab3.log
cat gpurun_out/lab3.log
