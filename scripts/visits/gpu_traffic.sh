#!/bin/bash
# (a) steady-state DRAM traffic per launch: three metrics = one pass, no replay, caches left alone (the persisting L2 window stays warm)
# (b) ncu --set full of three consecutive sweeps of the timed region (the dominant variant k_sweep<1,1,0>)
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:"k_sweep|k_update_vars" -s 40 -c 48 --csv --log-file gpurun_out/traffic_steady.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
tail -4 gpurun_out/traffic_steady.csv | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_update_vars" -s 30 -c 6 -f -o gpurun_out/prof_sweep python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
