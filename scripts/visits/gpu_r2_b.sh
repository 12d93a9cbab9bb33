#!/bin/bash
# round 2, visit B: metric-exchange test, profile of the TMA sweep kernel, scheduling / L2-promotion variants, e2e breakdown
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_group_tests.log
for v in "GBP_TILE_QUEUE=0" "GBP_TMA_L2PROMO=0" "GBP_TMA_L2PROMO=3" "GBP_L2_PERSIST=0"; do
  env $v timeout 300 python bench.py --steps 110 --warmup 11 --no-cpu-baseline > gpurun_out/r2b_bench_$v.json 2> gpurun_out/r2b_bench_$v.err
  echo "$v: $(python scripts/show_bench.py gpurun_out/r2b_bench_$v.json | cut -c1-160)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 44 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "k_sweep|k_update" gpurun_out/r2b_launches.csv | tail -24 | awk -F'","' '{print $5, $NF}' | tr -d '"' | sort | uniq -c | sort -rn | head -30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_update_vars" -s 20 -c 2 -f -o gpurun_out/r2b_prof_sweep python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 300 python scripts/e2e_breakdown.py 2>&1 | tail -40 | tee gpurun_out/r2b_e2e_breakdown.log
ls -la gpurun_out | tail -4
