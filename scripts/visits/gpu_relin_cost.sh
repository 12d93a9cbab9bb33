#!/bin/bash
# per-launch durations of the two-pass sweep flavour around a lock-step relinearisation sweep, and of k_relinearise_all
mkdir -p gpurun_out
cat > /tmp/two_pass.py <<PY
import sys
sys.path.insert(0, ".")
import bench
from gbp_poplar_b200 import GBPEngine, default_opts
bal, setup = bench.build_problem()
eng = GBPEngine(setup.problem, default_opts(relin_mode=2, use_cuda_graph=0))
bench.ba_preroll(eng)
eng.iterate(30)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_two_pass.csv python /tmp/two_pass.py > gpurun_out/ncu_two_pass.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/launches_two_pass.csv")))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        h = r; start = i + 1; break
idx = {n: i for i, n in enumerate(h)}
out = [(r[idx["Kernel Name"]][:28], float(r[idx["Metric Value"]]) / 1e3) for r in rows[start:] if len(r) >= len(h)]
print([o for o in out if "relinearise_all" in o[0]])
print(out[-100:])
PY
