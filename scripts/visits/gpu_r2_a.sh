#!/bin/bash
# round 2, visit A: group tests, the GPU suite on the TMA kernel, A/B of the two sweep kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_group_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_group_tests.log
timeout 1500 python -m pytest tests -q -x -m gpu --deselect tests/test_group_gpu.py 2>&1 | tail -12 | tee gpurun_out/r2_gpu_tests.log
for k in cpasync tma; do
  GBP_SWEEP=$k timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2a_bench_$k.json 2> gpurun_out/r2a_bench_$k.err
  python scripts/show_bench.py gpurun_out/r2a_bench_$k.json; tail -2 gpurun_out/r2a_bench_$k.err
done
