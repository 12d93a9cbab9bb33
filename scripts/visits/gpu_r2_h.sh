#!/bin/bash
# round 2, visit H: problem setup on the device (gbp_setup.cu), device helper unit tests: parity suite, bench, e2e breakdown
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_gpu_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_k20.json 2> gpurun_out/r2h_bench_k20.err
echo "K=20: $(python scripts/show_bench.py gpurun_out/r2h_bench_k20.json | cut -c1-170)"; tail -2 gpurun_out/r2h_bench_k20.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2h_bench_k20.json') if l.startswith('{')][-1]); print('e2e', d['e2e']); print('parity', d['parity_n']['status'])"
timeout 300 python scripts/e2e_breakdown.py 2>&1 | grep -v "iterate: \|^\[gbp shard\]" | tail -24 | tee gpurun_out/r2h_e2e_breakdown.log
