"""Generate golden output vectors from the REFERENCE's own codelet sources.

Runs oracle/_ref/libgbp_ref.so (= /root/reference/ba/gbp_codelets.cpp +
matlib.cpp + bafuncs.cpp compiled unmodified behind oracle/shim) in the
authoring container and freezes small fixtures that pin the oracle port and
the CUDA path on machines where /root/reference does not exist:

  golden_helpers.npz   known answers of inv6x6 / inv3x3 / hfunc+Jac on random inputs
  golden_runs.npz      final beliefs of short runs on the three sequences
  golden_runs.json     per-sweep metrics and SHA-256 of every tensor at checkpoints

Usage (authoring container only):  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
import oracle_lib  # noqa: E402
from gbp_poplar_b200 import MODE_SLAM  # noqa: E402
from gbp_poplar_b200.engine import TENSOR_NAMES  # noqa: E402

KIND = "reference"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sha_all(eng):
    """SHA-256 of every tensor; for the tensors of common.LOWER_ONLY also of their lower-triangle form."""
    out = {t: sha(eng.get_tensor(t)) for t in TENSOR_NAMES}
    for t in common.LOWER_ONLY:
        out[t + ":lower"] = sha(common.canon(t, eng.get_tensor(t)))
    return out


def helpers():
    rng = np.random.default_rng(20201017)
    A6, A3, X, P = [], [], [], []
    for _ in range(64):
        M = rng.normal(size=(6, 8))
        A6.append((M @ M.T + np.diag(rng.uniform(0.1, 10, 6))).astype(np.float32) * np.float32(10 ** rng.uniform(-2, 4)))
        M = rng.normal(size=(3, 5))
        A3.append((M @ M.T + np.diag(rng.uniform(0.1, 10, 3))).astype(np.float32) * np.float32(10 ** rng.uniform(-2, 4)))
        x = rng.normal(size=6) * [0.5, 0.5, 0.5, 0.4, 0.4, 0.4]
        X.append(x.astype(np.float32))
        P.append((rng.normal(size=3) * [1, 1, 0.5] + [0, 0, 3]).astype(np.float32))
    K = np.array([517.306408, 0, 318.64304, 0, 516.469215, 255.313989, 0, 0, 1], np.float32)
    A6, A3, X, P = map(np.array, (A6, A3, X, P))
    out = dict(A6=A6, A3=A3, X=X, P=P, K=K)
    out["inv6"] = np.array([oracle_lib.inv6x6(a, KIND) for a in A6])
    out["inv3"] = np.array([oracle_lib.inv3x3(a, KIND) for a in A3])
    pr = [oracle_lib.project(x, p, K, KIND) for x, p in zip(X, P)]
    out["hx"] = np.array([r[0] for r in pr])
    out["Jkf"] = np.array([r[1] for r in pr])
    out["Jlmk"] = np.array([r[2] for r in pr])
    np.savez_compressed(os.path.join(HERE, "golden_helpers.npz"), **out)


def runs():
    arrays, meta = {}, {}
    for name, n_sweeps, checkpoints in (("fr2robot2", 40, (0, 17, 39)), ("fr1xyz", 60, (0, 17, 19, 59)),
                                        ("fr1desk", 30, (0, 29))):
        for order in (0, 1):
            st = common.make_setup(name)
            eng = oracle_lib.OracleEngine(st.problem, kind=KIND)
            eng.set_reduce_order(order)
            key = f"{name}_ba_order{order}"
            m = {"init": eng.eval(), "sweeps": [], "sha": {}}
            m["sha"]["init"] = sha_all(eng)
            for it in range(n_sweeps):
                common.ba_schedule_step(eng, it)
                m["sweeps"].append(eng.eval())
                if it in checkpoints:
                    m["sha"][str(it)] = sha_all(eng)
            b = eng.get_beliefs()
            for t in ("cam_beliefs_eta", "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda"):
                arrays[f"{key}_{t}"] = b[t]
            meta[key] = m
            eng.close()
    # SLAM schedule on fr2robot2 (ba/slam.cpp loop), 25 sweeps per keyframe
    for order in (0, 1):
        st = common.make_setup("fr2robot2", mode=MODE_SLAM)
        eng = oracle_lib.OracleEngine(st.problem, kind=KIND)
        eng.set_reduce_order(order)
        new_lmks = []
        finals = common.slam_run(eng, st, 25, on_kf=lambda dc, n: new_lmks.append(n))
        key = f"fr2robot2_slam25_order{order}"
        meta[key] = {"finals": finals, "new_lmks": new_lmks,
                     "sha": {"final": sha_all(eng)}}
        b = eng.get_beliefs()
        for t in ("cam_beliefs_eta", "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda"):
            arrays[f"{key}_{t}"] = b[t]
        eng.close()
    # ---- long horizons of BASELINE.json configs 1-3 (CUDA summation order), pinned without the oracle at test time
    long_runs = {}
    # config 1: fr1xyz, ./ba default = 1500 sweeps
    st = common.make_setup("fr1xyz")
    eng = oracle_lib.OracleEngine(st.problem, kind=KIND)
    eng.set_reduce_order(1)
    series = []
    for it in range(1500):
        common.ba_schedule_step(eng, it)
        series.append(eng.eval()["reproj_mean"])
    arrays["long_fr1xyz_ba1500_reproj"] = np.array(series, np.float64)
    long_runs["fr1xyz_ba1500"] = {"sha": {t: sha(eng.get_tensor(t)) for t in ("cam_beliefs_eta", "cam_beliefs_lambda",
                                                                                 "lmk_beliefs_eta", "lmk_beliefs_lambda",
                                                                                 "damping_count", "robust_flag")}}
    eng.close()
    # config 2: fr1desk "to convergence": the reference diverges after ~250-330 sweeps; the series pins
    # the descent, the minimum window and the stop rule (first sweep above 2x the running minimum)
    st = common.make_setup("fr1desk")
    eng = oracle_lib.OracleEngine(st.problem, kind=KIND)
    eng.set_reduce_order(1)
    series = []
    for it in range(360):
        common.ba_schedule_step(eng, it)
        series.append(eng.eval()["reproj_mean"])
    arrays["long_fr1desk_ba360_reproj"] = np.array(series, np.float64)
    eng.close()
    # config 3: fr2robot2, ./slam default = 700 sweeps between keyframes (13 299 sweeps)
    st = common.make_setup("fr2robot2", mode=MODE_SLAM)
    eng = oracle_lib.OracleEngine(st.problem, kind=KIND)
    eng.set_reduce_order(1)
    finals = common.slam_run(eng, st, 700)
    arrays["long_fr2robot2_slam700_reproj"] = np.array([f["reproj_mean"] for f in finals], np.float64)
    long_runs["fr2robot2_slam700"] = {"sha": {t: sha(eng.get_tensor(t)) for t in ("cam_beliefs_eta", "cam_beliefs_lambda",
                                                                                     "lmk_beliefs_eta", "lmk_beliefs_lambda",
                                                                                     "damping_count", "robust_flag")}}
    eng.close()
    meta["long_runs"] = long_runs
    np.savez_compressed(os.path.join(HERE, "golden_runs.npz"), **arrays)
    with open(os.path.join(HERE, "golden_runs.json"), "w") as f:
        json.dump(meta, f, indent=0)


if __name__ == "__main__":
    assert oracle_lib.available("reference"), "needs oracle/_ref (make -C oracle ref in the authoring container)"
    helpers()
    runs()
    print("golden vectors written to", HERE)
