#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -6
for m in 0 1; do
  GBP_SKIP_UPPER=$m timeout 600 python bench.py --no-cpu-baseline --steps 300 > gpurun_out/bench_skip$m.json 2> gpurun_out/bench_skip$m.err
  echo "SKIP_UPPER=$m"; python scripts/show_bench.py gpurun_out/bench_skip$m.json | cut -c1-170; tail -2 gpurun_out/bench_skip$m.err
done
