#!/bin/bash
# the other mxv consumers: unchanged reference drivers, unbuffered output
mkdir -p gpurun_out
G=/tmp/chesapeake.mtx
cp tests/golden/chesapeake.mtx $G
{
for d in gmis gcc ggc glgc gdiameter; do
  [ -x build/dropin/$d ] || { echo "### $d not built"; continue; }
  rm -f /tmp/.chesapeake.mtx.*
  echo "### $d"
  timeout 120 stdbuf -o0 -e0 build/dropin/$d --mxvmode 0 --niter 1 --timing 0 --directed 2 $G 2>&1 | tail -40
  echo "### exit ${PIPESTATUS[0]}"
done
echo "### gmis under compute-sanitizer"
rm -f /tmp/.chesapeake.mtx.*
timeout 300 compute-sanitizer --tool memcheck build/dropin/gmis --mxvmode 0 --niter 1 --timing 0 --directed 2 $G 2>&1 | grep -v "^\[" | head -60
} > gpurun_out/extra.log 2>&1
cat gpurun_out/extra.log | cut -c1-300 | tail -150
