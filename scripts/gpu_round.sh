#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full ncu capture of the factor kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -45 gpurun_out/launches.csv | cut -c1-220
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 20 -c 2 -f -o gpurun_out/prof_sweep python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
