#!/bin/bash
# bench line + launch list only (kernel tuning loop)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "free_running_ba_bit_exact or synthetic_medium" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python scripts/show_bench.py gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 24 --csv --log-file gpurun_out/launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "k_sweep|k_update" gpurun_out/launches.csv | tail -12 | awk -F'","' '{print $5, $NF}' | tr -d '"'
