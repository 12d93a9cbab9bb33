#!/bin/bash
# round 2: GPU parity suite (one GPU), log kept
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m pytest tests/test_group_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_group_tests.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_group_gpu.py 2>&1 | tail -12 | tee gpurun_out/r2_gpu_tests.log
