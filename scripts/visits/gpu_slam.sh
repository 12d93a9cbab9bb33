#!/bin/bash
# SLAM keyframe insertion: device path vs the reference's host round trip (tests + timings)
mkdir -p gpurun_out /tmp/seq
python -m pytest tests -m gpu -x -q -k "slam or golden" 2>&1 | tail -5
python - <<PY
import sys; sys.path.insert(0,"tests")
import common
from gbp_poplar_b200 import BALProblem
common.load_sequence("fr2robot2").save("/tmp/seq/fr2robot2.txt")
BALProblem.synthetic(1000, 100000, 10.0, seed=1234).save("/tmp/seq/synth4.txt")
PY
for hk in 0 1; do
  ./gbp_poplar_b200/bin/slam --bal_file /tmp/seq/fr2robot2.txt --host_keyframes $hk > gpurun_out/slam_fr2robot2_700_hk$hk.log 2>&1
  grep -E "Timing|Keyframe insertions" gpurun_out/slam_fr2robot2_700_hk$hk.log
  tail -4 gpurun_out/slam_fr2robot2_700_hk$hk.log | grep Iters
done
for hk in 0 1; do
  ./gbp_poplar_b200/bin/slam --bal_file /tmp/seq/synth4.txt --iters_between_kfs 2 --host_keyframes $hk > gpurun_out/slam_synth4_hk$hk.log 2>&1
  grep -E "Timing|Keyframe insertions" gpurun_out/slam_synth4_hk$hk.log
  grep Iters gpurun_out/slam_synth4_hk$hk.log | tail -1
done
cmp <(grep -v -E "Timing|Keyframe insertions" gpurun_out/slam_synth4_hk0.log) <(grep -v -E "Timing|Keyframe insertions" gpurun_out/slam_synth4_hk1.log) && echo "synth4 logs identical"
