#!/bin/bash
# round 2 multi-GPU visit (gpurun --gpus N): sharded parity tests on real peers + the weak-scaling bench at N
N=${1:-2}; TESTS=${2:-yes}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$TESTS" = yes ]; then timeout 1200 python -m pytest tests/test_multigpu.py tests/test_group_gpu.py -q -x -m gpu 2>&1 | tail -60 | tee gpurun_out/r2_multigpu_tests_n$N.log | tail -6; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python scripts/show_bench.py gpurun_out/r2_bench_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r2_bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1])
print("e2e", d["e2e"]); print("config", d["config"]["parallelism"], d["config"]["rank0_shard"], "init_s", d["config"]["init_s"])
print("parity_n", json.dumps(d.get("parity_n"))[:600]); print("per_rank_kernel_us", d["roofline"].get("per_rank_kernel_us"))
PY
