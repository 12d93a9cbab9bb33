#!/bin/bash
# round-2 closing visit on ONE GPU: parity suite (log kept), smoke, both bench arms at the driver's and at the default
# step counts, launch list, steady-state DRAM traffic, full ncu capture of the sweep kernel, sanitizers, layout micro-benchmark
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8 | tee gpurun_out/r2_gpu_tests_1gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_k20.json 2> gpurun_out/r2_bench_k20.err; echo "K=20: $(python scripts/show_bench.py gpurun_out/r2_bench_k20.json | cut -c1-260)"; tail -2 gpurun_out/r2_bench_k20.err
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "default: $(python scripts/show_bench.py gpurun_out/r2_bench_n1.json | cut -c1-260)"; tail -2 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 44 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "k_sweep|k_update" gpurun_out/r2_launches.csv | tail -8 | awk -F'","' '{print $5, $NF}' | tr -d '"'
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:"k_sweep|k_update_vars" -s 40 -c 48 --csv --log-file gpurun_out/r2_traffic_steady.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
tail -3 gpurun_out/r2_traffic_steady.csv | cut -c1-220
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_tma|k_update_vars" -s 30 -c 4 -f -o gpurun_out/r2_prof_sweep python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/r2_prof_sweep.ncu-rep
bash scripts/gpu_sanitize.sh
./build/bin/membench 2>&1 | tee gpurun_out/r2_membench.log
timeout 600 python scripts/size_sweep.py 2>&1 | tail -5 | tee gpurun_out/r2_size_sweep.log
