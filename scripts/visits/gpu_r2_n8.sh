#!/bin/bash
# round 2 closing visit on 8 GPUs: sharded parity tests between processes, the weak-scaling bench, configs[4] exactly
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu --tb=short 2>&1 | tail -40 > gpurun_out/r2_multigpu_tests_n$N.log; tail -3 gpurun_out/r2_multigpu_tests_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python scripts/show_bench.py gpurun_out/r2_bench_n$N.json | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload config5 --steps 200 --warmup 10 > gpurun_out/r2_bench_config5_n$N.json 2> gpurun_out/r2_bench_config5_n$N.err
python scripts/show_bench.py gpurun_out/r2_bench_config5_n$N.json | cut -c1-330
python - <<PY
import json
for f in ("gpurun_out/r2_bench_n$N.json", "gpurun_out/r2_bench_config5_n$N.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["ms_per_step"], "parity", (d.get("parity_n") or {}).get("status"), d["config"].get("rank0_shard"), "init_s", d["config"].get("init_s"))
        print("  per_rank_kernel_us", d["roofline"].get("per_rank_kernel_us"))
    except Exception as e: print(f, "unreadable", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r2_bench_n$N.err gpurun_out/r2_bench_config5_n$N.err | tail -4
