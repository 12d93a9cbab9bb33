"""Single-GPU sweep time on the N=2-sized graph (2 k cameras / 200 k landmarks, 6.5 % robust factors)."""
import sys, time
sys.path.insert(0, ".")
import bench
from gbp_poplar_b200 import GBPEngine, default_opts
bal, setup = bench.build_problem(scale=2)
eng = GBPEngine(setup.problem, default_opts())
bench.ba_preroll(eng)
eng.iterate(20)
eng.iterate(110)
ms, _ = eng.last_timing()
eng.set_profile(True); eng.iterate(110); a, b = eng.last_kernel_times(); eng.set_profile(False)
print("us/sweep %.1f  k_sweep %.1f  k_vars %.1f  robust %d" % (ms / 110 * 1e3, a / 110 * 1e3, b / 110 * 1e3, eng.eval()["n_robust"]), flush=True)
