#!/bin/bash
# round-end visit on one GPU (short form): smoke, both bench arms, launch list, steady-state traffic, full ncu of three sweeps
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python scripts/show_bench.py gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:"k_sweep|k_update_vars" -s 40 -c 48 --csv --log-file gpurun_out/traffic_steady.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 44 --csv --log-file gpurun_out/launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
GBP_CUDA_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stats.csv python scripts/metric_launches.py > /dev/null 2>&1
grep -E "k_" gpurun_out/launches_stats.csv | tail -4 | awk -F'","' '{print $5, $NF}' | tr -d '"'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_update_vars" -s 30 -c 4 -f -o gpurun_out/prof_sweep python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
