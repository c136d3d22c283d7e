#!/bin/bash
# Other consumers of the mxv path ("next"): the UNCHANGED
# reference drivers gmis, glgc, gdiameter compile against this backend too
# (__graft_entry__.build_dropin builds them best-effort).  Not part of the r01
# parity set: this script is the first thing to run for that row on a GPU box.
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
G=tests/golden/chesapeake.mtx
{
for d in gmis glgc gdiameter; do
  [ -x build/dropin/$d ] || { echo "### $d not built"; continue; }
  echo "### build/dropin/$d --mxvmode 0 --niter 1 --timing 0 --directed 2 $G"
  timeout 300 build/dropin/$d --mxvmode 0 --niter 1 --timing 0 --directed 2 $G
  echo "### exit $?"
done
} > "$OUT/dropin_extra.log" 2>&1
grep -c "^CORRECT" "$OUT/dropin_extra.log"; grep -n "INCORRECT\|Error\|exit [1-9]" "$OUT/dropin_extra.log" | head
