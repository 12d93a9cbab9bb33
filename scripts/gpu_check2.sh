#!/bin/bash
# two-GPU visit: all GPU tests (single + sharded + CLI), then the N=1 and N=2 bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python scripts/show_bench.py gpurun_out/bench_n1.json | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python scripts/show_bench.py gpurun_out/bench_n2.json | cut -c1-200
