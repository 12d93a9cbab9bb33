"""Host-side setup: BAL loader, priors / scaling / flags, SLAM bookkeeping, synthetic generator."""
import math
import os

import numpy as np
import pytest

import common
from gbp_poplar_b200 import BALProblem, Setup, cli_options, MODE_BA, MODE_SLAM


def test_save_load_roundtrip(tmp_path):
    bal = BALProblem.synthetic(12, 150, 6.0, seed=3)
    path = tmp_path / "synth.txt"
    bal.save(path)
    again = BALProblem.load(path)
    assert (again.n_keyframes, again.n_points, again.n_edges) == (bal.n_keyframes, bal.n_points, bal.n_edges)
    assert np.array_equal(again.camera_index, bal.camera_index)
    assert np.array_equal(again.point_index, bal.point_index)
    assert np.allclose(again.observations, bal.observations, rtol=0, atol=1e-3)
    assert np.array_equal(again.parameters, bal.parameters)
    # the text layout is the reference's (sequences/README.md:5-16): header, intrinsics, E records, 6C+3L values
    lines = open(path).read().split("\n")
    assert lines[0].split() == [str(bal.n_keyframes), str(bal.n_points), str(bal.n_edges)]
    assert len([ln for ln in lines if ln.strip()]) == 2 + bal.n_edges + 6 * bal.n_keyframes + 3 * bal.n_points


def test_missing_file_raises():
    with pytest.raises(FileNotFoundError, match="unable to open file"):
        BALProblem.load("/nonexistent/file.txt")


def test_truncated_file_does_not_crash(tmp_path, capfd):
    p = tmp_path / "short.txt"
    p.write_text("2 3 4\n500 500 320 240\n0 0 1.0 2.0\n")
    bal = BALProblem.load(p)
    assert (bal.n_keyframes, bal.n_points, bal.n_edges) == (2, 3, 4)
    assert "Invalid UW data file." in capfd.readouterr().out


def test_ba_setup_matches_reference_formulas():
    st = common.make_setup("fr1xyz")
    p = st.problem
    assert (p.n_keyframes, p.n_points, p.n_edges) == (42, 2194, 12908)
    K = st.K
    assert K[0] == np.float32(517.306408) and K[4] == np.float32(516.469215)
    assert K[2] == np.float32(318.64304) and K[5] == np.float32(255.313989) and K[8] == 1
    assert np.all(st.array("meas_variances") == 4.0)
    assert np.all(st.array("active_flag") == 1)
    assert np.all(st.array("cam_weaken_flag") == 5) and np.all(st.array("lmk_weaken_flag") == 5)
    assert np.all(st.array("damping_count") == -15) and np.all(st.array("damping") == 0)
    lam = st.array("cam_priors_lambda").reshape(42, 6, 6)
    eta = st.array("cam_priors_eta").reshape(42, 6)
    for c in (0, 5, 41):
        d = np.diag(lam[c])
        assert np.all(d == d[0]) and d[0] > 0 and np.count_nonzero(lam[c]) == 6
        mean = common.load_sequence("fr1xyz").parameters[6 * c:6 * c + 6].astype(np.float32)
        assert np.allclose(eta[c], mean * d[0], rtol=1e-6)
    s = st.array("cam_scaling")
    # ba/ba.cpp:560-572
    assert s[2] == pytest.approx(math.exp(-2 / 5 * math.log(100.0)), rel=1e-6)
    assert s[0] == pytest.approx(math.exp(-1 / 5 * math.log(float(lam[0, 0, 0]) * 0.01 ** 2)), rel=1e-6)
    assert np.all(st.array("lmk_scaling") == np.float32(s[2]))
    # after `steps` weakenings the first cameras' prior std is first_cam_prior_std
    assert float(lam[0, 0, 0]) * float(s[0]) ** 5 == pytest.approx(1 / 0.01 ** 2, rel=1e-4)


def test_options_change_setup():
    st = common.make_setup("fr2robot2", reproj_meas_var=9.0, iters_before_damping=7, steps=3.0)
    assert np.all(st.array("meas_variances") == 9.0)
    assert np.all(st.array("damping_count") == -7)
    assert np.all(st.array("cam_weaken_flag") == 3)


def test_avdepth_and_noise_options_run():
    a = common.make_setup("fr2robot2")
    b = common.make_setup("fr2robot2", av_depth_on=1)
    assert not np.array_equal(a.array("lmk_priors_eta"), b.array("lmk_priors_eta"))
    c1 = common.make_setup("fr2robot2", transnoise=0.05, rotnoise=1.0, lmktrans_noise=0.02, noise_seed=5)
    c2 = common.make_setup("fr2robot2", transnoise=0.05, rotnoise=1.0, lmktrans_noise=0.02, noise_seed=5)
    assert np.array_equal(c1.array("cam_priors_eta")[:12], a.array("cam_priors_eta")[:12])  # first two cameras anchored
    assert not np.array_equal(c1.array("cam_priors_eta"), a.array("cam_priors_eta"))
    assert np.array_equal(c1.array("lmk_priors_eta"), c2.array("lmk_priors_eta"))


def test_slam_flag_schedule():
    st = common.make_setup("fr2robot2", mode=MODE_SLAM)
    cam = st.array("cam_ids")
    act = st.array("active_flag")
    assert np.array_equal(act == 1, cam <= 1)  # create_flags: cameras 0 and 1 (ba/dataio.cpp:455-475)
    assert list(st.array("cam_weaken_flag")[:3]) == [5, 5, 0]
    seen = np.zeros(st.problem.n_points, bool)
    seen[st.array("lmk_ids")[cam <= 1]] = True
    assert np.array_equal(st.array("lmk_weaken_flag") == 5, seen)


def test_setup_rejects_bad_indices():
    with pytest.raises(RuntimeError):
        bal = BALProblem.from_arrays([500, 500, 320, 240], [0, 3], [0, 0], np.zeros(4), np.zeros(12), np.zeros(3))
        Setup(bal)


def test_synthetic_generator():
    a = BALProblem.synthetic(40, 1500, 8.0, seed=11)
    b = BALProblem.synthetic(40, 1500, 8.0, seed=11)
    c = BALProblem.synthetic(40, 1500, 8.0, seed=12)
    assert np.array_equal(a.observations, b.observations) and np.array_equal(a.parameters, b.parameters)
    assert not np.array_equal(a.observations[:100], c.observations[:100])
    ci, li = a.camera_index, a.point_index
    assert np.all(np.diff(ci.astype(np.int64)) >= 0), "edges must be camera-sorted (quirk Q7)"
    deg = np.bincount(li, minlength=a.n_points)
    assert deg.min() >= 2
    assert 5.0 * a.n_points < a.n_edges < 9.0 * a.n_points
    ob = a.observations.reshape(-1, 2)
    assert ob[:, 0].min() > 10 and ob[:, 0].max() < 630 and ob[:, 1].min() > 10 and ob[:, 1].max() < 470
    # no duplicate (camera, landmark) pairs
    assert len(set(zip(ci.tolist(), li.tolist()))) == a.n_edges
    # rotations stay away from the |w| = 0 singularity of Jac (bafuncs.cpp:197-204)
    w = a.parameters[:6 * a.n_keyframes].reshape(-1, 6)[:, 3:]
    assert np.linalg.norm(w, axis=1).min() > 1e-3
