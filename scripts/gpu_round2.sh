#!/bin/bash
# quick GPU visit: parity tests + bench line (+ optional variants given as GBP_CUDA_LIB paths)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -8
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; python scripts/show_bench.py gpurun_out/bench_main.json; tail -3 gpurun_out/bench_main.err
for lib in "$@"; do
  echo "== variant $lib"
  GBP_CUDA_LIB=$lib timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$(basename $lib .so).json 2>gpurun_out/err.txt; python scripts/show_bench.py gpurun_out/bench_$(basename $lib .so).json; tail -3 gpurun_out/err.txt
done
