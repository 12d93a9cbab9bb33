"""Is a shard's graph itself harder than the one-GPU graph?  One GPU: the one-GPU bench graph, then the sub-problems of
rank 0 / rank 1 of the 2-rank partition and of an interior rank of the 8-rank partition run as PLAIN handles (no
exchange at all): ms per sweep and the per-kernel event times."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gbp_poplar_b200 import GBPEngine  # noqa: E402
from gbp_poplar_b200.host import Shard  # noqa: E402


def run(problem, label):
    eng = GBPEngine(problem)
    bench.ba_preroll(eng)
    eng.iterate(20)
    eng.iterate(220)
    ms = eng.last_timing()[0] / 220
    eng.set_profile(True)
    eng.iterate(44)
    tf, tv = eng.last_sweep_times()
    eng.set_profile(False)
    plain = tf < 1.4 * tf.min()      # sweeps without relinearisation
    print(f"{label}: {problem.n_keyframes} cameras {problem.n_points} landmarks {problem.n_edges} factors: "
          f"{ms * 1e3:.1f} us/sweep; event-timed k_sweep {tf[plain].mean() * 1e3:.1f} us (plain sweeps) / {tf.mean() * 1e3:.1f} (all), "
          f"k_update_vars {tv.mean() * 1e3:.1f} us", flush=True)
    eng.close()


bal, st = bench.build_problem(1)
run(st.problem, "one-GPU graph")
for world, ranks in ((2, (0, 1)), (8, (3,))):
    bal, st = bench.build_problem(world)
    for r in ranks:
        sh = Shard(st.problem, world, r, owner=st)
        run(sh.problem, f"rank {r} of {world} as a plain handle")
