#!/bin/bash
# perf visit: GPU tests, bench line, launch list, full ncu of k_sweep / k_update_vars
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -12
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python scripts/show_bench.py gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 24 --csv --log-file gpurun_out/launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "k_sweep|k_update" gpurun_out/launches.csv | tail -24 | awk -F'","' '{print $5, $NF}' | tr -d '"'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_update_vars" -s 20 -c 2 -f -o gpurun_out/prof_sweep python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
