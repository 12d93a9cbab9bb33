"""Sweep time against graph size on one GPU (0.5 / 1 / 2 / 4 M factors): what happens past the 48 MB persisting-L2 window
that holds the landmark-bound messages of config 4.  Prints ns per factor and sweep for each size."""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gbp_poplar_b200 import BALProblem, GBPEngine, Setup, default_opts  # noqa: E402

rows = []
for cams, lmks in ((500, 50000), (1000, 100000), (2000, 200000), (4000, 400000)):
    st = Setup(BALProblem.synthetic(cams, lmks, 10.5, 1234))
    E = st.problem.n_edges
    eng = GBPEngine(st.problem, default_opts())
    bench.ba_preroll(eng)
    eng.iterate(11)
    eng.iterate(110)
    ms, _ = eng.last_timing()
    eng.set_profile(True)
    eng.iterate(110)
    f, v = eng.last_kernel_times()
    eng.close()
    row = {"cameras": cams, "landmarks": lmks, "factors": E, "us_per_sweep": ms / 110 * 1e3, "ns_per_factor_sweep": ms / 110 * 1e6 / E,
           "k_sweep_us": f / 110 * 1e3, "k_update_vars_us": v / 110 * 1e3, "landmark_messages_MB": E * 48 / 1e6,
           "G_factor_updates_per_s": E * 110 / (ms / 1e3) / 1e9}
    rows.append(row)
    print(json.dumps(row), flush=True)
