#!/bin/bash
mkdir -p gpurun_out
G=/tmp/chesapeake.mtx
cp tests/golden/chesapeake.mtx $G
{
for d in gmis gcc; do
  rm -f /tmp/.chesapeake.mtx.*
  echo "### $d"
  timeout 300 cuda-gdb -batch -ex run -ex bt --args build/dropin/$d --mxvmode 0 --niter 1 --timing 0 --directed 2 $G 2>&1 | grep -v "^\[New\|^\[Thread\|^warning\|Detaching\|^\[Switching" | tail -45
done
} > gpurun_out/gdb.log 2>&1
cat gpurun_out/gdb.log | cut -c1-400
