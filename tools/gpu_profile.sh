#!/bin/bash
# One GPU session of evidence: benches for the four algorithms, ncu launch lists
# and --set full captures of the hot kernels.  Usage: tools/gpu_profile.sh <tag>
TAG=${1:-cur}
OUT=gpurun_out/r01
mkdir -p $OUT
python bench.py --steps 10 --warmup 3 > $OUT/bench_bfs24_$TAG.json 2> $OUT/bench_bfs24_$TAG.err
python bench.py --algo sssp --steps 3 --warmup 1 > $OUT/bench_sssp24_$TAG.json 2> $OUT/bench_sssp24_$TAG.err
python bench.py --algo sssp --scale 22 --steps 5 --warmup 2 > $OUT/bench_sssp22_$TAG.json 2> $OUT/bench_sssp22_$TAG.err
python bench.py --algo pr --scale 22 --steps 3 --warmup 1 --no-cpu-baseline > $OUT/bench_pr22_$TAG.json 2> $OUT/bench_pr22_$TAG.err
python bench.py --algo tc --scale 22 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_tc22_$TAG.json 2> $OUT/bench_tc22_$TAG.err
python bench.py --algo tc --scale 18 --steps 2 --warmup 1 > $OUT/bench_tc18_$TAG.json 2> $OUT/bench_tc18_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference_bfs24_$TAG.json 2> $OUT/bench_reference_bfs24_$TAG.err
# launch lists (device time per launch; cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_bfs24_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_sssp22_$TAG.csv python bench.py --algo sssp --scale 22 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_l2.log 2>&1
# full captures of the hot kernels
ncu --set full --clock-control none --import-source on -k regex:spmvMergeKernelT -s 2 -c 1 \
    -o $OUT/prof_merge_$TAG -f python bench.py --algo sssp --scale 22 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spmvMaskedOrPullBitsKernel|spmspvPushKernel" -s 7 -c 7 \
    -o $OUT/prof_bfs_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmspvPushKernel -s 4 -c 3 \
    -o $OUT/prof_ssspush_$TAG -f python bench.py --algo sssp --scale 24 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spgemmMasked" -c 2 \
    -o $OUT/prof_tc_$TAG -f python bench.py --algo tc --scale 20 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_f4.log 2>&1
for f in bfs24 sssp24 sssp22 pr22 tc22 tc18 reference_bfs24; do echo "== $f"; cut -c1-400 $OUT/bench_${f}_$TAG.json; tail -2 $OUT/bench_${f}_$TAG.err; done
ls -la $OUT | tail -15
