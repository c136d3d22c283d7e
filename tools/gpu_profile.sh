#!/bin/bash
# One GPU session of evidence: benches for the four algorithms, ncu launch lists
# and --set full captures of the hot kernels.  Usage: tools/gpu_profile.sh <tag>
TAG=${1:-cur}
OUT=gpurun_out/r01
mkdir -p $OUT
# gpurun copies back at most 64 MiB: captures are summarised ON THE BOX
# (tools/summarize_ncu.py) and only the two main .ncu-rep files are kept.
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.txt 2>&1 || { tail -20 $OUT/pytest_gpu_$TAG.txt; exit 1; }
tail -2 $OUT/pytest_gpu_$TAG.txt
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
S=$OUT/summaries_$TAG
mkdir -p $S
python tools/summarize_ncu.py launches $OUT/launches_bfs24_$TAG.csv > $S/launches_bfs_rmat24.txt
python tools/summarize_ncu.py launches $OUT/launches_sssp22_$TAG.csv > $S/launches_sssp_rmat22.txt
python tools/summarize_ncu.py full $OUT/prof_merge_$TAG.ncu-rep > $S/ncu_spmv_merge_rmat22.txt
python tools/summarize_ncu.py full $OUT/prof_bfs_$TAG.ncu-rep > $S/ncu_bfs_kernels_rmat24.txt
python tools/summarize_ncu.py full $OUT/prof_ssspush_$TAG.ncu-rep > $S/ncu_sssp_push_rmat24.txt
python tools/summarize_ncu.py full $OUT/prof_tc_$TAG.ncu-rep > $S/ncu_tc_rmat20.txt
cp profiles/traffic.json $S/traffic_before.json 2>/dev/null
python tools/summarize_ncu.py traffic $OUT/prof_merge_$TAG.ncu-rep sssp:22:0 spmvMergeKernelT
python tools/summarize_ncu.py traffic $OUT/prof_merge_$TAG.ncu-rep pr:22:0 spmvMergeKernelT
python tools/summarize_ncu.py traffic $OUT/prof_bfs_$TAG.ncu-rep bfs:24:1 spmvMaskedOrPullBitsKernel
python tools/summarize_ncu.py traffic $OUT/prof_ssspush_$TAG.ncu-rep sssp:24:2 spmspvPushKernel
python tools/summarize_ncu.py traffic $OUT/prof_tc_$TAG.ncu-rep tc:20:3 spgemmMasked
cp profiles/traffic.json $S/traffic.json
ncu -i $OUT/prof_ssspush_$TAG.ncu-rep --page source --csv > $S/ncu_sssp_push_source.csv 2>/dev/null
rm -f $OUT/prof_ssspush_$TAG.ncu-rep $OUT/prof_tc_$TAG.ncu-rep $OUT/prof_bfs_$TAG.ncu-rep
du -sh gpurun_out
for f in bfs24 sssp24 sssp22 pr22 tc22 tc18 reference_bfs24; do echo "== $f"; cut -c1-400 $OUT/bench_${f}_$TAG.json; tail -2 $OUT/bench_${f}_$TAG.err; done
ls -la $OUT | tail -15
