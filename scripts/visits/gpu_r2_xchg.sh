#!/bin/bash
# what the exchange blocks cost the belief update at N=2 (results are wrong in modes 2 and 3: timing only), and the same
# two shards driven from ONE process (GBPGroup: direct peer pointers, no CUDA IPC)
mkdir -p gpurun_out
for v in 0 3 2; do
  GBP_XCHG_DEBUG=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_xchg_$v.json 2> gpurun_out/r2_xchg_$v.err
  echo "GBP_XCHG_DEBUG=$v: $(python scripts/show_bench.py gpurun_out/r2_xchg_$v.json | cut -c1-140)"
done
timeout 300 python scripts/exchange_timeline.py 2 2>&1 | grep "per sweep"
GBP_XCHG_DEBUG=3 timeout 300 python scripts/exchange_timeline.py 2 2>&1 | grep "per sweep"
