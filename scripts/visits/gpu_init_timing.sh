#!/bin/bash
# phases of gbp_cuda_init on the bench workload (stderr lines of GBP_INIT_TIMING=1)
GBP_INIT_TIMING=1 python - <<PY
import sys, time
sys.path.insert(0, ".")
import bench
from gbp_poplar_b200 import GBPEngine, default_opts
bal, setup = bench.build_problem()
for i in range(3):
    t0 = time.time(); e = GBPEngine(setup.problem, default_opts()); t1 = time.time()
    e.iterate(3)
    t2 = time.time(); b = e.get_beliefs(); t3 = time.time()
    print(f"init {1e3*(t1-t0):.1f} ms  get_beliefs {1e3*(t3-t2):.1f} ms", flush=True)
    e.close()
PY
