#!/bin/bash
# N=2 bench with the handle arena from the stream-ordered pool (1) or from cudaMalloc (0)
for pool in 0 1; do
GBP_ARENA_POOL=$pool timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
echo "GBP_ARENA_POOL=$pool"; python scripts/show_bench.py gpurun_out/ab.json | cut -c1-140
done
