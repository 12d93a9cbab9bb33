#!/bin/bash
# multi-GPU visit (gpurun --gpus N): sharded parity tests + bench at N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 600 python -m pytest tests/test_multigpu.py -q -x -m gpu 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python scripts/show_bench.py gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
