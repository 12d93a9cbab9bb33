#!/bin/bash
# round 2, visit G: k_update_vars with bulk-staged camera partials / two-round-trip landmark blocks; 8 vs 10 warps per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests.log
timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2g_bench_tw8.json 2> gpurun_out/r2g_bench_tw8.err
echo "tw8: $(python scripts/show_bench.py gpurun_out/r2g_bench_tw8.json | cut -c1-150)"; tail -1 gpurun_out/r2g_bench_tw8.err
for v in 1 2; do
  GBP_UV_DEBUG=$v timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2g_bench_uv$v.json 2> gpurun_out/r2g_bench_uv$v.err
  echo "GBP_UV_DEBUG=$v: $(python scripts/show_bench.py gpurun_out/r2g_bench_uv$v.json | cut -c1-150)"
done
export GBP_CUDA_LIB=$PWD/build/variants/libgbp_tw10.so
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_group_gpu.py -q -x -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2g_bench_tw10.json 2> gpurun_out/r2g_bench_tw10.err
echo "tw10: $(python scripts/show_bench.py gpurun_out/r2g_bench_tw10.json | cut -c1-150)"; tail -1 gpurun_out/r2g_bench_tw10.err
