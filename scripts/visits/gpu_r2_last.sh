#!/bin/bash
# round-2 last visit on ONE GPU with the final build: parity suite (log kept), smoke, bench at the default and the driver's
# step counts, launch list, fast-math memcheck
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8 | tee gpurun_out/r2_gpu_tests_1gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
timeout 300 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "default: $(python scripts/show_bench.py gpurun_out/r2_bench_n1.json | cut -c1-260)"; tail -2 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_k20.json 2> gpurun_out/r2_bench_k20.err; echo "K=20: $(python scripts/show_bench.py gpurun_out/r2_bench_k20.json | cut -c1-260)"; tail -2 gpurun_out/r2_bench_k20.err
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_run.py fast > gpurun_out/r2_sanitizer_memcheck_fast_math.log 2>&1
echo "== memcheck, fast-math build: rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r2_sanitizer_memcheck_fast_math.log | tail -1)"; grep "fast-math" gpurun_out/r2_sanitizer_memcheck_fast_math.log | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 44 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "k_sweep|k_update" gpurun_out/r2_launches.csv | tail -6 | awk -F'","' '{print $5, $NF}' | tr -d '"'
