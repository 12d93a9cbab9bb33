"""Where the time of a short ba-style job goes (host clock per C-ABI call): init phases (GBP_INIT_TIMING=1 prints
them on stderr), the first and the steady-state gbp_cuda_iterate(1, stats), weaken_priors, get_beliefs."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GBP_INIT_TIMING", "1")
import bench  # noqa: E402
from gbp_poplar_b200 import GBPEngine, default_opts  # noqa: E402

bal, setup = bench.build_problem(1)
for rep in range(3):
    t = [time.perf_counter()]
    eng = GBPEngine(setup.problem, default_opts())
    t.append(time.perf_counter())
    its = []
    for it in range(20):
        a = time.perf_counter()
        if (it + 1) % 2 == 0 and it < 10:
            eng.weaken_priors()
        b = time.perf_counter()
        eng.iterate(1, stats=True)
        c = time.perf_counter()
        its.append((b - a, c - b))
    t.append(time.perf_counter())
    bel = eng.get_beliefs()
    t.append(time.perf_counter())
    eng.close()
    t.append(time.perf_counter())
    print(f"rep {rep}: init {1e3*(t[1]-t[0]):.2f} ms  loop {1e3*(t[2]-t[1]):.2f} ms  read {1e3*(t[3]-t[2]):.2f} ms  close {1e3*(t[4]-t[3]):.2f} ms")
    print("   iterate ms:", " ".join(f"{1e3*x[1]:.2f}" for x in its))
    print("   weaken ms :", " ".join(f"{1e3*x[0]:.2f}" for x in its[:10]))
