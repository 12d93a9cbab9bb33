"""Pin the CPU oracle: the restated port must reproduce, bit for bit, the golden
vectors generated from the reference's own codelet sources (tests/golden/make_golden.py),
and -- where oracle/_ref exists -- the reference build itself on fresh inputs."""
import hashlib
import json
import os

import numpy as np
import pytest

import common
import oracle_lib
from gbp_poplar_b200 import MODE_SLAM
from gbp_poplar_b200.engine import TENSOR_NAMES

G = common.GOLDEN


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def runs_meta():
    with open(os.path.join(G, "golden_runs.json")) as f:
        return json.load(f)


def test_helpers_match_golden():
    z = np.load(os.path.join(G, "golden_helpers.npz"))
    for i in range(len(z["A6"])):
        assert oracle_lib.inv6x6(z["A6"][i]).tobytes() == z["inv6"][i].tobytes()
        assert oracle_lib.inv3x3(z["A3"][i]).tobytes() == z["inv3"][i].tobytes()
        hx, jk, jl = oracle_lib.project(z["X"][i], z["P"][i], z["K"])
        assert hx.tobytes() == z["hx"][i].tobytes()
        assert jk.tobytes() == z["Jkf"][i].tobytes()
        assert jl.tobytes() == z["Jlmk"][i].tobytes()


def test_inverse_is_an_inverse():
    z = np.load(os.path.join(G, "golden_helpers.npz"))
    for A in z["A6"][:16]:
        Ai = oracle_lib.inv6x6(A).astype(np.float64)
        assert np.abs(Ai @ A.astype(np.float64) - np.eye(6)).max() < 5e-3
    for A in z["A3"][:16]:
        Ai = oracle_lib.inv3x3(A).astype(np.float64)
        assert np.abs(Ai @ A.astype(np.float64) - np.eye(3)).max() < 1e-3


@pytest.mark.parametrize("name,n_sweeps", [("fr2robot2", 40), ("fr1xyz", 60), ("fr1desk", 30)])
@pytest.mark.parametrize("order", [0, 1])
def test_ba_runs_match_golden(runs_meta, name, n_sweeps, order):
    m = runs_meta[f"{name}_ba_order{order}"]
    st = common.make_setup(name)
    eng = oracle_lib.OracleEngine(st.problem, kind="port")
    eng.set_reduce_order(order)
    for t in TENSOR_NAMES:
        assert sha(eng.get_tensor(t)) == m["sha"]["init"][t], t
    assert eng.eval() == pytest.approx(m["init"])
    for it in range(n_sweeps):
        common.ba_schedule_step(eng, it)
        if str(it) in m["sha"]:
            for t in TENSOR_NAMES:
                assert sha(eng.get_tensor(t)) == m["sha"][str(it)][t], (it, t)
            assert eng.eval() == pytest.approx(m["sweeps"][it])
    z = np.load(os.path.join(G, "golden_runs.npz"))
    b = eng.get_beliefs()
    for t in ("cam_beliefs_eta", "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda"):
        assert b[t].tobytes() == z[f"{name}_ba_order{order}_{t}"].tobytes()


def test_slam_run_matches_golden(runs_meta):
    m = runs_meta["fr2robot2_slam25_order0"]
    st = common.make_setup("fr2robot2", mode=MODE_SLAM)
    eng = oracle_lib.OracleEngine(st.problem, kind="port")
    new_lmks = []
    finals = common.slam_run(eng, st, 25, on_kf=lambda dc, n: new_lmks.append(n))
    assert new_lmks == m["new_lmks"]
    # 18 insertions with these new-landmark counts
    assert new_lmks == [31, 19, 33, 17, 29, 19, 22, 19, 25, 40, 54, 34, 59, 102, 73, 41, 19, 0]
    for a, b in zip(finals, m["finals"]):
        assert a == pytest.approx(b)
    for t in TENSOR_NAMES:
        assert sha(eng.get_tensor(t)) == m["sha"]["final"][t], t


def test_known_answers_from_survey():
    """SURVEY.md section 8c known answers (independent probe of the reference codelets)."""
    st = common.make_setup("fr1xyz")
    eng = oracle_lib.OracleEngine(st.problem, kind="port")
    assert eng.eval()["reproj_mean"] == pytest.approx(199.1097, rel=1e-5)
    assert (eng.max_nkfedges, eng.max_nlmkedges) == (419, 30)
    common.ba_schedule_step(eng, 0)
    s = eng.eval()
    assert s["reproj_mean"] == pytest.approx(146.0789, rel=1e-5)
    assert s["n_robust"] == 12903
    st = common.make_setup("fr1desk")
    eng = oracle_lib.OracleEngine(st.problem, kind="port")
    assert eng.eval()["reproj_mean"] == pytest.approx(209.693, rel=1e-5)
    assert (eng.max_nkfedges, eng.max_nlmkedges) == (383, 46)
    st = common.make_setup("fr2robot2", mode=MODE_SLAM)
    eng = oracle_lib.OracleEngine(st.problem, kind="port")
    assert eng.eval()["reproj_mean"] == pytest.approx(32.7526, rel=1e-5)
    assert (eng.max_nkfedges, eng.max_nlmkedges) == (280, 15)


def test_thread_count_does_not_change_results():
    st = common.make_setup("fr2robot2")
    outs = []
    for threads in (1, 3):
        eng = oracle_lib.OracleEngine(st.problem, kind="port", threads=threads)
        common.run_ba(eng, 25)
        outs.append(eng.snapshot())
    for t in TENSOR_NAMES:
        assert outs[0][t].tobytes() == outs[1][t].tobytes(), t


@pytest.mark.skipif(not oracle_lib.available("reference"), reason="oracle/_ref not built (no /root/reference)")
def test_port_equals_reference_build_live():
    rng = np.random.default_rng(7)
    K = np.array([517.3, 0, 318.6, 0, 516.5, 255.3, 0, 0, 1], np.float32)
    for _ in range(200):
        M = rng.normal(size=(6, 7))
        A = (M @ M.T + np.eye(6)).astype(np.float32)
        assert oracle_lib.inv6x6(A, "port").tobytes() == oracle_lib.inv6x6(A, "reference").tobytes()
        x = (rng.normal(size=6) * 0.5).astype(np.float32)
        p = (rng.normal(size=3) + [0, 0, 3]).astype(np.float32)
        for u, v in zip(oracle_lib.project(x, p, K, "port"), oracle_lib.project(x, p, K, "reference")):
            assert u.tobytes() == v.tobytes()
    st = common.make_setup("fr2robot2")
    a = oracle_lib.OracleEngine(st.problem, kind="port")
    b = oracle_lib.OracleEngine(st.problem, kind="reference")
    common.run_ba(a, 30)
    common.run_ba(b, 30)
    for t in TENSOR_NAMES:
        assert a.get_tensor(t).tobytes() == b.get_tensor(t).tobytes(), t


def test_codelet_level_sequence_equals_iterate():
    st = common.make_setup("fr2robot2")
    a = oracle_lib.OracleEngine(st.problem, kind="port")
    b = oracle_lib.OracleEngine(st.problem, kind="port")
    for it in range(22):
        common.ba_schedule_step(a, it)
        if (it + 1) % 2 == 0 and it < 10:
            b.weaken_prior_vertices()
            b.update_beliefs()
        b.prep_messages()
        b.compute_messages()
        b.update_beliefs()
        b.commit_messages()
    for t in TENSOR_NAMES:
        assert a.get_tensor(t).tobytes() == b.get_tensor(t).tobytes(), t
