#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5 | cut -c1-200
timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print('bench default: ms %.3f value %.0f e2e %.0f launches %d parity %s clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['parity_vs_cpu_reference'], d['clocks']))"
