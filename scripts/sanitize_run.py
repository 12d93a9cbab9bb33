"""A short run of every kernel family of the library for compute-sanitizer (scripts/gpu_sanitize.sh): init (device setup
kernels), the ba.cpp schedule with prior weakening, block calls with and without per-sweep metrics, both sweep flavours,
the SLAM keyframe insertion on the device, tensor read-backs.  Small graph (fr2robot2: 20 cameras / 3.5 k factors) so the
instrumented run takes a minute."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import common  # noqa: E402
from gbp_poplar_b200 import GBPEngine, MODE_SLAM, default_opts  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "main"
if what == "fast":   # the opt-in contracted-FMA build of the sweep kernel, on its own
    import faulthandler
    faulthandler.enable()
    st = common.make_setup("fr2robot2")   # (kept alive: gbp_problem points into its arrays)
    fast = GBPEngine(st.problem, default_opts(fast_math=1))
    print("fast-math: handle built", flush=True)
    common.run_ba(fast, 12)
    print("fast-math: schedule done", flush=True)
    fast.iterate(5)
    print("fast-math", fast.eval()["reproj_mean"], flush=True)
    fast.close()
    sys.exit(0)
from gbp_poplar_b200 import _capi  # noqa: E402
for relin_mode in (1, 2):
    st = common.make_setup("fr2robot2")
    eng = GBPEngine(st.problem, default_opts(relin_mode=relin_mode))
    common.run_ba(eng, 12)
    eng.iterate(9)
    stats = eng.iterate(3, stats=True)
    eng.get_beliefs()
    eng.get_tensor("cam_messages_lambda")
    print("relin_mode", relin_mode, "reproj", stats[-1]["reproj_mean"], flush=True)
    eng.close()
    _capi.load_library().gbp_cuda_release_cached_memory()   # the instrumented run is short of device memory
st = common.make_setup("fr2robot2", mode=MODE_SLAM)
eng = GBPEngine(st.problem)
finals = common.slam_run(eng, st, 6, device_kf=True, stats_every_kf=False)
print("slam final", finals[-1]["reproj_mean"], flush=True)
eng.close()
