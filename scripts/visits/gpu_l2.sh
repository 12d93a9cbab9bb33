#!/bin/bash
# L2 persisting window for the re-used buffers: window size (from the start of the landmark-bound messages) x set-aside size
mkdir -p gpurun_out
run() {
  GBP_INIT_TIMING=1 GBP_L2_PERSIST=$1 GBP_L2_WINDOW_MB=$2 GBP_L2_SETASIDE_MB=$3 timeout 600 python bench.py --no-cpu-baseline --steps 300 > gpurun_out/bench_l2.json 2> gpurun_out/bench_l2.err
  echo "mode=$1 window=$2 setaside=$3 $(grep 'L2 window' gpurun_out/bench_l2.err | head -1)"; python scripts/show_bench.py gpurun_out/bench_l2.json | cut -c1-150
}
run 1 1000 1000
run 1 1000 40
run 1 1000 32
run 1 1000 24
run 2 56 1000
run 2 65 1000
run 2 1000 64
run 1 1000 1000
