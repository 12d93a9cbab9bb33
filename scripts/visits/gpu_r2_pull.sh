#!/bin/bash
# boundary exchange: partial sums read from the peers (default) against stored into the peers (GBP_XCHG_PUSH=1), and where
# the finish blocks sit among the landmark blocks (GBP_FINISH_AT, percent), at N ranks
N=${1:-2}
mkdir -p gpurun_out
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_pull_${label}_n$N.json 2> gpurun_out/r2_pull_${label}_n$N.err
  echo "$label: $(python scripts/show_bench.py gpurun_out/r2_pull_${label}_n$N.json | cut -c1-150)"
}
run pull_at0 GBP_FINISH_AT=0
run pull_at40 GBP_FINISH_AT=40
run pull_at80 GBP_FINISH_AT=80
run pull_at100 GBP_FINISH_AT=100
run push_at0 GBP_XCHG_PUSH=1 GBP_FINISH_AT=0
run pull_at0_again GBP_FINISH_AT=0
