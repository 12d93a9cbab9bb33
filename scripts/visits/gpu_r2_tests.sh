#!/bin/bash
# round 2: the whole GPU parity suite on one GPU, log kept
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1800 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -25 | tee gpurun_out/r2_gpu_tests.log
