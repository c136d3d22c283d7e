#!/bin/bash
# Copies the small Matrix Market fixtures the reference's own tests and examples
# run on (reference data/small/, used by test/gvxm.cu:239-300, test/greduce.cu:63-75
# and BASELINE.json configs[0]) into tests/golden/.  /root/reference does not
# exist on the GPU box, so the data files are committed next to this script.
set -e
SRC=${1:-/root/reference/data/small}
DST=$(dirname "$0")
for f in chesapeake.mtx test_cc.mtx test_sgm.mtx test_bc.mtx; do
  cp "$SRC/$f" "$DST/$f"
done
