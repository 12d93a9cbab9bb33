#!/bin/bash
# where the finish blocks of the boundary exchange sit among the landmark blocks (GBP_FINISH_AT, percent): bench at N ranks
N=${1:-2}; shift
mkdir -p gpurun_out
for v in "$@"; do
  GBP_FINISH_AT=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_finish_at_${v}_n$N.json 2> gpurun_out/r2_finish_at_${v}_n$N.err
  echo "GBP_FINISH_AT=$v: $(python scripts/show_bench.py gpurun_out/r2_finish_at_${v}_n$N.json | cut -c1-150)"
done
