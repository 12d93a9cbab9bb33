#!/bin/bash
# round 2, visit D: layout micro-benchmark, what bounds k_update_vars (camera / landmark halves skipped in turn), e2e breakdown
mkdir -p gpurun_out
./build/bin/membench 2>&1 | tee gpurun_out/r2d_membench.log
./build/bin/membench 66288 2>&1 | tee -a gpurun_out/r2d_membench.log
for v in 0 1 2 3; do
  GBP_UV_DEBUG=$v timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2d_bench_uv$v.json 2> gpurun_out/r2d_bench_uv$v.err
  echo "GBP_UV_DEBUG=$v: $(python scripts/show_bench.py gpurun_out/r2d_bench_uv$v.json | cut -c1-150)"; tail -1 gpurun_out/r2d_bench_uv$v.err
done
timeout 300 python scripts/e2e_breakdown.py 2>&1 | grep -v "iterate: \|^\[gbp shard\]" | tail -24 | tee gpurun_out/r2d_e2e_breakdown.log
