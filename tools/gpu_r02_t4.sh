#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r02_extra.sh > /dev/null 2>&1
grep -n "###\|CORRECT\|INCORRECT\|not implemented\|rror" gpurun_out/extra.log | cut -c1-160 | head -60
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t4_pytest.log
tail -25 gpurun_out/t4_pytest.log
