#!/bin/bash
# small-graph latency (configs 1-3) through the drop-in CLI tools, on frozen copies of the reference sequences
mkdir -p gpurun_out /tmp/seq
python - <<PY
import sys; sys.path.insert(0,"tests")
import common
for n in ("fr1xyz","fr1desk","fr2robot2"):
    common.load_sequence(n).save(f"/tmp/seq/{n}.txt")
PY
for i in 1 2; do ./gbp_poplar_b200/bin/ba --bal_file /tmp/seq/fr1xyz.txt > gpurun_out/ba_fr1xyz.log 2>&1; done
grep -E "Initial Reprojection|^Iter 1499|Timing report" gpurun_out/ba_fr1xyz.log
./gbp_poplar_b200/bin/ba --bal_file /tmp/seq/fr1desk.txt --n_iters 360 > gpurun_out/ba_fr1desk.log 2>&1; grep -E "^Iter 359|Timing report" gpurun_out/ba_fr1desk.log
./gbp_poplar_b200/bin/slam --bal_file /tmp/seq/fr2robot2.txt --iters_between_kfs 100 > gpurun_out/slam_fr2robot2_100.log 2>&1; tail -3 gpurun_out/slam_fr2robot2_100.log | grep -E "Iters|Timing"
./gbp_poplar_b200/bin/slam --bal_file /tmp/seq/fr2robot2.txt > gpurun_out/slam_fr2robot2_700.log 2>&1; grep -E "Timing" gpurun_out/slam_fr2robot2_700.log
