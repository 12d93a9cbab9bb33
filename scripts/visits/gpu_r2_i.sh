#!/bin/bash
# round 2, visit I: fast-math tests + the whole GPU suite, bench with the fma_mode object
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_gpu_tests.log
timeout 400 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python scripts/show_bench.py gpurun_out/r2i_bench.json | cut -c1-330; tail -2 gpurun_out/r2i_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2i_bench.json') if l.startswith('{')][-1]); print('fma', d['fma_mode']); print('classes', d['roofline'].get('classes'))"
