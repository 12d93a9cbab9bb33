#!/bin/bash
# bench several builds of the library: bash scripts/visits/gpu_variants.sh lib1.so lib2.so ...
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  GBP_CUDA_LIB=$lib timeout 600 python bench.py --no-cpu-baseline --steps 300 > gpurun_out/bench_$(basename $lib .so).json 2>gpurun_out/err.txt; python scripts/show_bench.py gpurun_out/bench_$(basename $lib .so).json | cut -c1-170; tail -3 gpurun_out/err.txt
done
