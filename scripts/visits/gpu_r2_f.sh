#!/bin/bash
# round 2, visit F: camera beliefs finished inside the sweep (tickets): parity suite, then fused on/off x warps per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests.log
for f in 1 0; do
  GBP_FUSE_CAM=$f timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2f_bench_tw8_f$f.json 2> gpurun_out/r2f_bench_tw8_f$f.err
  echo "tw8 fuse=$f: $(python scripts/show_bench.py gpurun_out/r2f_bench_tw8_f$f.json | cut -c1-150)"; tail -1 gpurun_out/r2f_bench_tw8_f$f.err
done
for tw in 9 10 11; do
  export GBP_CUDA_LIB=$PWD/build/variants/libgbp_tw$tw.so
  timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2f_bench_tw$tw.json 2> gpurun_out/r2f_bench_tw$tw.err
  echo "tw$tw: $(python scripts/show_bench.py gpurun_out/r2f_bench_tw$tw.json | cut -c1-150)"; tail -1 gpurun_out/r2f_bench_tw$tw.err
  unset GBP_CUDA_LIB
done
GBP_CUDA_LIB=$PWD/build/variants/libgbp_tw10.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_group_gpu.py -q -x -m gpu 2>&1 | tail -3
