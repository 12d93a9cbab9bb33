"""Five sweeps with the per-sweep metric on the config-4 graph (for an ncu launch list of the stats path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gbp_poplar_b200 import GBPEngine  # noqa: E402

bal, setup = bench.build_problem()
eng = GBPEngine(setup.problem)
for _ in range(5):
    print(eng.iterate(1, stats=True)[0])
