"""Ad-hoc GPU check: CUDA path vs the CPU oracle on a reference sequence."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import common  # noqa: E402
import oracle_lib  # noqa: E402
from gbp_poplar_b200 import GBPEngine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "fr1xyz"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
st = common.make_setup(name)
kind = "reference" if oracle_lib.available("reference") else "port"
ora = oracle_lib.OracleEngine(st.problem, kind=kind)
ora.set_reduce_order(1)
gpu = GBPEngine(st.problem)
print("oracle", kind, "init", ora.eval())
print("gpu init", gpu.eval())
names = list(common.BLOCK_DIMS)


def cmp(tag):
    a, b = gpu.snapshot(names + ["damping", "damping_count", "robust_flag", "mu", "dmu"]), ora.snapshot(names + ["damping", "damping_count", "robust_flag", "mu", "dmu"])
    worst = {k: float(common.block_rel_err(a[k], b[k], d).max()) for k, d in common.BLOCK_DIMS.items()}
    exact = {k: a[k].tobytes() == b[k].tobytes() for k in a}
    print(tag, "max block rel err:", {k: f"{v:.2e}" for k, v in worst.items()})
    print(tag, "bit-identical:", [k for k, v in exact.items() if v], "| differing:", [k for k, v in exact.items() if not v])


cmp("after init")
for it in range(n):
    for eng in (ora, gpu):
        common.ba_schedule_step(eng, it)
    if it in (0, 1, 4, 16, 17, 18, 19, n - 1):
        cmp(f"sweep {it}")
        print("   oracle", ora.eval())
        print("   gpu   ", gpu.eval())
t = time.time()
gpu.iterate(200)
ms, k = gpu.last_timing()
print(f"200 sweeps: device {ms:.3f} ms ({ms/200*1000:.1f} us/sweep), {k} kernels, wall {time.time()-t:.3f}s")
