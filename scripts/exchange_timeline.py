"""Where the time of a sharded belief update goes (diagnostic build: make EXTRA=-DGBP_DEBUG_TS, GBP_DEBUG_TS=1): all ranks
in one process (GBPGroup, one shard per GPU), 32 sweeps, then per rank the mean offsets from the first block of
k_update_vars to: last push block announced / first and last finish block past the wait / last finish block done /
last landmark block done / last camera block done."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GBP_DEBUG_TS"] = "1"
import bench  # noqa: E402
from gbp_poplar_b200 import GBPGroup, _capi, default_opts  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bal, setup = bench.build_problem(world)
import torch  # noqa: E402
ndev = torch.cuda.device_count()
grp = GBPGroup(setup.problem, world, devices=[r % ndev for r in range(world)], opts=default_opts())
bench.ba_preroll(grp)
grp.iterate(40)
lib = _capi.load_library()
for r in grp.ranks:   # re-arm
    buf = (C.c_uint64 * 256)()
    lib.gbp_cuda_debug_timestamps(r.handle, buf, 256)
grp.iterate(32)
ms, _ = grp.last_timing()
print(f"world {world}: {ms / 32 * 1e3:.1f} us per sweep")
names = ["push announced", "first finish past wait", "last finish past wait", "last finish done", "last landmark done", "last camera done"]
for rank, r in enumerate(grp.ranks):
    buf = (C.c_uint64 * 256)()
    n = lib.gbp_cuda_debug_timestamps(r.handle, buf, 256)
    t = np.array(buf[:n], dtype=np.uint64).reshape(-1, 8).astype(np.float64)
    ok = (t[:, 0] < 1.8e19) & (t[:, 5] > 0)
    d = (t[ok][:, 1:7] - t[ok][:, :1]) / 1e3
    d[d < 0] = np.nan
    print(f"rank {rank} ({ok.sum()} exchanges), us after the first block:", ", ".join(f"{nm} {v:.1f}" for nm, v in zip(names, np.nanmean(d, axis=0))))
grp.close()
