#!/bin/bash
# round-end visit on one GPU: all GPU tests, smoke, both bench arms, launch list, full ncu of the two sweep kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python scripts/show_bench.py gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 44 --csv --log-file gpurun_out/launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "k_sweep|k_update" gpurun_out/launches.csv | tail -6 | awk -F'","' '{print $5, $NF}' | tr -d '"'
GBP_CUDA_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_stats.csv python scripts/metric_launches.py > /dev/null 2>&1
grep -E "k_" gpurun_out/launches_stats.csv | tail -5 | awk -F'","' '{print $5, $NF}' | tr -d '"'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_update_vars" -s 20 -c 2 -f -o gpurun_out/prof_sweep python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
# the same capture with the caches left as they are between replays: DRAM traffic with the landmark-bound messages resident in the
# persisting L2 window (the default capture flushes L2 before every replay)
timeout 900 ncu --set full --clock-control none --cache-control none -k regex:"k_sweep|k_update_vars" -s 20 -c 2 -f -o gpurun_out/prof_sweep_warm python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full_warm.log 2>&1
ls -la gpurun_out | tail -4
