#!/bin/bash
# r02 call 1: hub SpMV lab (sanitizer at small scale, then RMAT-22/24) + relabel measurement
mkdir -p gpurun_out
L=build/lab/spmv_lab
echo "== sanitizer scale 14 ==" > gpurun_out/lab1.log
timeout 300 compute-sanitizer --tool memcheck $L 14 16 1 >> gpurun_out/lab1.log 2>&1
echo "rc=$?" >> gpurun_out/lab1.log
echo "== scale 16 ==" >> gpurun_out/lab1.log
timeout 120 $L 16 16 3 >> gpurun_out/lab1.log 2>&1
echo "rc=$?" >> gpurun_out/lab1.log
echo "== scale 22 ==" >> gpurun_out/lab1.log
timeout 300 $L 22 16 5 >> gpurun_out/lab1.log 2>&1
echo "rc=$?" >> gpurun_out/lab1.log
echo "== scale 24 ==" >> gpurun_out/lab1.log
timeout 300 $L 24 16 3 >> gpurun_out/lab1.log 2>&1
echo "rc=$?" >> gpurun_out/lab1.log
echo "== relabel off ==" >> gpurun_out/lab1.log
timeout 300 python bench.py --algo sssp --scale 22 --steps 5 --warmup 3 --no-cpu-baseline >> gpurun_out/lab1.log 2>gpurun_out/lab1_err0.log
echo "== relabel on ==" >> gpurun_out/lab1.log
GB200_SPMV_RELABEL=1 timeout 300 python bench.py --algo sssp --scale 22 --steps 5 --warmup 3 --no-cpu-baseline >> gpurun_out/lab1.log 2>gpurun_out/lab1_err1.log
tail -c 6000 gpurun_out/lab1.log
