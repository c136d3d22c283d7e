#!/bin/bash
# r02 evidence in one GPU call: bench lines, reference arm, ncu launch lists, ncu
# --set full captures of the dominant kernels and their text summaries.
# Everything lands in gpurun_out/r02/ (copied to profiles/ by hand afterwards).
OUT=gpurun_out/r02
mkdir -p $OUT
run_bench() {   # name, args...
  local name=$1; shift
  timeout 900 python bench.py "$@" > $OUT/r02_bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r02_bench_$name.json")); r=d["roofline"]
    print("$name: ms/step %.3f  value %.0f  e2e %.0f  parity %s  dom %s  frac %.3f  share %.2f  launches/step %.1f"
          % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity_vs_cpu_reference"], r["kernel"][:28],
             r["frac"] or 0, r["share_of_step"], d["gpu_launches"]/d["steps"]))
except Exception as e:
    print("$name failed:", e); print(open("$OUT/bench_$name.err").read()[-800:])
PY
}
run_bench bfs_rmat24 --steps 20 --warmup 5
run_bench sssp_rmat22 --algo sssp --steps 10 --warmup 3
run_bench sssp_rmat24_pushpull --algo sssp --scale 24 --mxvmode 0 --steps 5 --warmup 3
run_bench pr_rmat22 --algo pr --steps 5 --warmup 3
run_bench tc_rmat22 --algo tc --steps 3 --warmup 3
run_bench tc_rmat20 --algo tc --scale 20 --steps 5 --warmup 3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference_arm_bfs_rmat24.json 2>/dev/null
timeout 600 python bench.py --impl reference --algo sssp --steps 2 --warmup 1 > $OUT/r02_bench_reference_arm_sssp_rmat22.json 2>/dev/null
# launch lists (cold cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv \
    --log-file $OUT/r02_launches_bfs_rmat24.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_ncu.py launches $OUT/r02_launches_bfs_rmat24.csv > $OUT/r02_launches_bfs_rmat24.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv \
    --log-file $OUT/r02_launches_sssp_rmat22.csv python bench.py --algo sssp --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_ncu.py launches $OUT/r02_launches_sssp_rmat22.csv > $OUT/r02_launches_sssp_rmat22.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/r02_launches_tc_rmat22.csv python bench.py --algo tc --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_ncu.py launches $OUT/r02_launches_tc_rmat22.csv > $OUT/r02_launches_tc_rmat22.txt 2>&1
# full captures of the dominant kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bfsFusedKernel -s 2 -c 1 -f \
    -o $OUT/prof_bfs_fused python bench.py --steps 2 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_ncu.py full $OUT/prof_bfs_fused.ncu-rep > $OUT/r02_ncu_bfs_fused_rmat24.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmvHubKernel -s 2 -c 1 -f \
    -o $OUT/prof_spmv_hub python bench.py --algo sssp --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_ncu.py full $OUT/prof_spmv_hub.ncu-rep > $OUT/r02_ncu_spmv_hub_rmat22.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spgemmHashKernel -s 6 -c 3 -f \
    -o $OUT/prof_tc python bench.py --algo tc --scale 20 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_ncu.py full $OUT/prof_tc.ncu-rep > $OUT/r02_ncu_tc_rmat20.txt 2>&1
python tools/summarize_ncu.py traffic $OUT/prof_bfs_fused.ncu-rep bfs:24:1 bfsFusedKernel
python tools/summarize_ncu.py traffic $OUT/prof_spmv_hub.ncu-rep sssp:22:0 spmvHubKernel
python tools/summarize_ncu.py traffic $OUT/prof_spmv_hub.ncu-rep pr:22:0 spmvHubKernel
cp profiles/traffic.json $OUT/traffic.json
# the reports themselves are too big to travel back (64 MiB limit): keep the text
rm -f $OUT/*.ncu-rep
ls -la $OUT | head -40
