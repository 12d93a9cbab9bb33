"""Block-relative error of the contracted-FMA build after ONE sweep from the reference's state (teacher-forced), per tensor:
max / p99.9 / p99 / fraction of blocks above 1e-4."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import common  # noqa: E402
import oracle_lib  # noqa: E402
from gbp_poplar_b200 import GBPEngine, default_opts  # noqa: E402

KIND = "reference" if oracle_lib.available("reference") else "port"
STATE = ["cam_messages_eta", "cam_messages_lambda", "lmk_messages_eta", "lmk_messages_lambda", "cam_beliefs_eta",
         "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda"]
for name, start in (("fr1xyz", 0), ("fr1xyz", 16), ("fr1xyz", 50), ("fr1xyz", 300), ("fr2robot2", 30), ("fr1desk", 40)):
    st = common.make_setup(name)
    ora = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ora.set_reduce_order(1)
    common.run_ba(ora, start)
    fast = GBPEngine(st.problem, default_opts(fast_math=1))
    fast.restore(ora.snapshot())
    ora.iterate(1)
    fast.iterate(1)
    for t in STATE:
        err = common.block_rel_err(common.canon(t, fast.get_tensor(t)), common.canon(t, ora.get_tensor(t)), common.BLOCK_DIMS[t])
        print(f"{name:10s} start {start:4d} {t:22s} max {err.max():.2e}  p99.9 {np.percentile(err, 99.9):.2e}  p99 {np.percentile(err, 99):.2e}  "
              f">1e-4: {(err > 1e-4).mean() * 100:.3f} %")
    fast.close()
