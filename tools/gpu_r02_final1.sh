#!/bin/bash
# full GPU suite, then the r02 evidence session
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | cut -c1-200
bash tools/gpu_r02_profile.sh
