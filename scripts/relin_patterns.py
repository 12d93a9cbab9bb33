"""Sweep time when relinearisations are scattered over sweeps instead of in lock step (config-4 graph)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gbp_poplar_b200 import GBPEngine  # noqa: E402

bal, setup = bench.build_problem()
E = setup.problem.n_edges
for mode in ("lockstep", "scattered"):
    eng = GBPEngine(setup.problem)
    if mode == "scattered":
        rng = np.random.default_rng(1)
        eng.set_tensor("damping_count", rng.integers(-25, -14, size=E).astype(np.int32))
    bench.ba_preroll(eng)
    eng.iterate(60)
    st = eng.iterate(44, stats=True)
    eng.iterate(110)
    ms, k = eng.last_timing()
    print(mode, "us/sweep %.1f" % (ms / 110 * 1e3), "relins per sweep:", [s["n_relins"] for s in st[::4]], flush=True)
