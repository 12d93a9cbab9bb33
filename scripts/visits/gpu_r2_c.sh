#!/bin/bash
# round 2, visit C: TMA kernel with the pipelined tile queue (8 warps, two buffers) against the single-buffered 10- and
# 12-warp builds: parity suite + bench for each; e2e breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests.log
timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2c_bench_tw8.json 2> gpurun_out/r2c_bench_tw8.err
echo "tw8: $(python scripts/show_bench.py gpurun_out/r2c_bench_tw8.json | cut -c1-150)"; tail -2 gpurun_out/r2c_bench_tw8.err
for tw in 10 12; do
  export GBP_CUDA_LIB=$PWD/build/variants/libgbp_tw$tw.so
  timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_group_gpu.py -q -x -m gpu 2>&1 | tail -3
  timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2c_bench_tw$tw.json 2> gpurun_out/r2c_bench_tw$tw.err
  echo "tw$tw: $(python scripts/show_bench.py gpurun_out/r2c_bench_tw$tw.json | cut -c1-150)"; tail -2 gpurun_out/r2c_bench_tw$tw.err
  unset GBP_CUDA_LIB
done
timeout 300 python scripts/e2e_breakdown.py 2>&1 | grep -v "^\[gbp shard\]" | tail -45 | tee gpurun_out/r2c_e2e_breakdown.log
