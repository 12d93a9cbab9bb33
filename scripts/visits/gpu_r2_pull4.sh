#!/bin/bash
# N=4 (or $1): the same shards with the partial sums stored into the peers (push) and read from them (pull)
N=${1:-4}
mkdir -p gpurun_out
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_pull_${label}_n$N.json 2> gpurun_out/r2_pull_${label}_n$N.err
  echo "$label: $(python scripts/show_bench.py gpurun_out/r2_pull_${label}_n$N.json | cut -c1-150)"
}
run push GBP_XCHG_PUSH=1
run pull GBP_XCHG_PUSH=0
run push_again GBP_XCHG_PUSH=1
