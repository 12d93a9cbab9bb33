#!/bin/bash
# round 2, visit E: the three staging flavours of the sweep kernel with the tile queue: cp.async / TMA (same structure) / TMA restructured
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests.log
for k in cpasync tma tma2; do
  GBP_SWEEP=$k timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2e_bench_$k.json 2> gpurun_out/r2e_bench_$k.err
  echo "$k: $(python scripts/show_bench.py gpurun_out/r2e_bench_$k.json | cut -c1-150)"; tail -1 gpurun_out/r2e_bench_$k.err
done
GBP_SWEEP=cpasync timeout 300 python -m pytest tests/test_parity_gpu.py -q -x -m gpu 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 22 -c 1 -f -o gpurun_out/r2e_prof_tma python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/r2e_prof_tma.ncu-rep
