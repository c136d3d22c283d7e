#!/bin/bash
mkdir -p gpurun_out
LAB_CEILING_ONLY=1 timeout 300 build/lab/spmv_lab 22 16 5 > gpurun_out/lab4.log 2>&1
echo "rc=$?" >> gpurun_out/lab4.log
cat gpurun_out/lab4.log
