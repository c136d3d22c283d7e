#!/bin/bash
# Runs the UNCHANGED reference drivers (compiled against this backend into
# build/dropin/) on the bundled small graph and on a generated R-MAT graph; each
# driver verifies itself against the reference's CPU implementation and prints
# CORRECT / INCORRECT.  Usage: tools/run_dropin_smoke.sh [out_dir] [sanitize]
OUT=${1:-gpurun_out}
SAN=${2:-}
mkdir -p "$OUT"
G=tests/golden/chesapeake.mtx
R=/tmp/rmat_s14.mtx
python tools/gen_rmat_mtx.py 14 16 $R
run() { echo "### $*"; if [ -n "$SAN" ]; then timeout 600 compute-sanitizer --error-exitcode 9 "$@"; else timeout 300 "$@"; fi; echo "### exit $?"; }
{
for graph in $G $R; do
  for mode in 0 1 2; do
    run build/dropin/gbfs --mxvmode $mode --struconly 1 --opreuse 1 --earlyexit 1 --niter 2 --timing 1 --directed 2 $graph
    run build/dropin/gbfs --mxvmode $mode --niter 1 --timing 0 --directed 2 $graph
    run build/dropin/gsssp --mxvmode $mode --niter 2 --timing 1 --directed 2 --seed 1 $graph
  done
  run build/dropin/gpr --mxvmode 0 --niter 2 --max_niter 10 --timing 1 --directed 2 $graph
  run build/dropin/gtc --mxvmode 0 --niter 1 --timing 1 --directed 2 $graph
done
} > "$OUT/dropin_smoke.log" 2>&1
grep -c "^CORRECT" "$OUT/dropin_smoke.log"; grep -n "INCORRECT\|errors occ\|Cuda error\|exit [1-9]\|ERROR SUMMARY" "$OUT/dropin_smoke.log" | head -40
