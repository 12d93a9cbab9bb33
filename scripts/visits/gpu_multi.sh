#!/bin/bash
# multi-GPU visit (gpurun --gpus N): [sharded parity tests +] bench at N
N=${1:-2}; TESTS=${2:-yes}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$TESTS" = yes ]; then timeout 900 python -m pytest tests/test_multigpu.py tests/test_cli.py -q -x -m gpu 2>&1 | tail -6; fi; if [ "$TESTS" = quick ]; then timeout 900 python -m pytest tests/test_multigpu.py -q -x -m gpu -k "synth" 2>&1 | tail -6; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python scripts/show_bench.py gpurun_out/bench_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
print("e2e", d["e2e"]); print("config", d["config"]["parallelism"], d["config"]["rank0_shard"], "init_s", d["config"]["init_s"])
PY
