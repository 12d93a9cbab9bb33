#!/bin/bash
# ncu --set full of the relinearisation kernels (k_relinearise_all at init)
mkdir -p gpurun_out
cat > /tmp/init_only.py <<PY
import sys
sys.path.insert(0, ".")
import bench
from gbp_poplar_b200 import GBPEngine, default_opts
bal, setup = bench.build_problem()
eng = GBPEngine(setup.problem, default_opts(use_cuda_graph=0))
eng.iterate(2)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_relinearise_all -c 1 -o gpurun_out/prof_relin_all -f python /tmp/init_only.py > gpurun_out/ncu_relin.log 2>&1
tail -3 gpurun_out/ncu_relin.log
