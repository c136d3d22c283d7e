#!/bin/bash
mkdir -p gpurun_out
L=build/lab/spmv_lab
echo "== scale 22 ==" > gpurun_out/lab7.log
for i in 1 2 3; do timeout 300 $L 22 16 3 2>&1 | grep -v "^stream\|^merge\|^ceiling" >> gpurun_out/lab7.log; done
echo "== scale 24 ==" >> gpurun_out/lab7.log
timeout 300 $L 24 16 3 2>&1 | grep -v "^stream\|^ceiling" >> gpurun_out/lab7.log
cat gpurun_out/lab7.log
