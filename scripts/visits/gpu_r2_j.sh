#!/bin/bash
# round 2, visit J: group tests with the shallower exchange chains, whole suite, fast-math error statistics
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --tb=short 2>&1 | tail -25 | tee gpurun_out/r2_gpu_tests.log
timeout 600 python scripts/fast_math_errors.py 2>&1 | tail -50 | tee gpurun_out/r2j_fast_math_errors.log
