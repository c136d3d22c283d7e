#!/bin/bash
# One GPU session: parity tests, headline benches, ncu launch list + full capture.
OUT=gpurun_out/r01
mkdir -p $OUT
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $OUT/pytest_gpu.txt
python bench.py --steps 10 --warmup 3 > $OUT/bench_bfs24.json 2> $OUT/bench_bfs24.err
python bench.py --algo sssp --steps 3 --warmup 1 > $OUT/bench_sssp24.json 2> $OUT/bench_sssp24.err
python bench.py --algo pr --scale 22 --steps 3 --warmup 1 --no-cpu-baseline > $OUT/bench_pr22.json 2> $OUT/bench_pr22.err
python bench.py --algo tc --scale 18 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_tc18.json 2> $OUT/bench_tc18.err
# launch list (device time per launch; cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_bfs24.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_bfs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $OUT/launches_sssp22.csv python bench.py --algo sssp --scale 22 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_sssp.log 2>&1
# full capture of the hot kernels
ncu --set full --clock-control none --import-source on -k regex:spmvMergeKernel -c 2 \
    -o $OUT/prof_merge python bench.py --algo sssp --scale 22 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_merge.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spmvMaskedOrPullKernel|spmspvPushKernel" -c 6 \
    -o $OUT/prof_bfs python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_bfsk.log 2>&1
ls -la $OUT
cat $OUT/pytest_gpu.txt
for f in bfs24 sssp24 pr22 tc18; do echo "== $f"; cut -c1-600 $OUT/bench_$f.json; tail -2 $OUT/bench_$f.err; done
