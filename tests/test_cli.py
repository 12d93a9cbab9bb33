"""The drop-in `ba` / `slam` command-line tools (gbp_poplar_b200/csrc/cli_main.cpp): flag surface
and error behaviour of ba/ba.cpp:394-487 on CPU; on a GPU the log lines and the numbers they carry."""
import os
import re
import subprocess

import numpy as np
import pytest

import common
from gbp_poplar_b200 import GBPEngine, MODE_SLAM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BA = os.path.join(ROOT, "gbp_poplar_b200", "bin", "ba")
SLAM = os.path.join(ROOT, "gbp_poplar_b200", "bin", "slam")
# ba/ba.cpp:400-465 ; slam replaces n_iters by iters_between_kfs (ba/slam.cpp:414-417)
COMMON_FLAGS = ["help", "bal_file", "profile", "ipus", "camspertile", "tn", "rn", "ltn", "avdepth_on", "avdepth",
                "reproj_meas_var", "prior_std_weaker_factor", "first_cam_prior_std", "steps", "undamped_start", "v"]


@pytest.fixture(scope="module", autouse=True)
def _built_tools():
    """The tools are build artefacts (gbp_poplar_b200/bin/, not in git): build them on demand."""
    if not (os.path.exists(BA) and os.path.exists(SLAM)):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "gbp_poplar_b200", "csrc"), "ba", "slam"])


def run(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, **kw)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def write_fixture(name, tmp_path):
    path = os.path.join(str(tmp_path), f"{name}.txt")
    common.load_sequence(name).save(path)
    return path


@pytest.mark.parametrize("exe,own", [(BA, "n_iters"), (SLAM, "iters_between_kfs")])
def test_help_lists_every_reference_flag(exe, own):
    r = run([exe, "--help"])
    assert r.returncode != 0          # the reference prints the options and then throws (ba.cpp:469-472)
    listed = re.findall(r"^\s+--(\w+)", r.stdout, flags=re.M)
    for f in COMMON_FLAGS + [own]:
        assert f in listed, f
    assert "(=1500)" in r.stdout or own != "n_iters"
    assert "(=700)" in r.stdout or own != "iters_between_kfs"
    for default in ("(=4)", "(=100)", "(=5)", "(=15)"):
        assert default in r.stdout


def test_argument_errors():
    r = run([BA])
    assert r.returncode == 1 and "'--bal_file' is required but missing" in r.stderr
    r = run([BA, "--bal_file", "/nonexistent/file.txt"])
    assert r.returncode == 1 and "ERROR: unable to open file /nonexistent/file.txt" in r.stderr   # ba.cpp:484-487
    r = run([BA, "--bal_file", "x", "--bogus", "1"])
    assert r.returncode == 1 and "unrecognised option '--bogus'" in r.stderr
    r = run([BA, "--bal_file", "x", "--v", "maybe"])
    assert r.returncode == 1 and "invalid" in r.stderr


@pytest.mark.skipif(_has_gpu(), reason="a GPU is present")
def test_no_device_exits_like_the_reference(tmp_path):
    path = write_fixture("fr2robot2", tmp_path)
    r = run([BA, "--bal_file", path, "--n_iters=3"])
    assert r.returncode != 0 and "Could not find a device" in r.stdout     # ba.cpp:652-655
    assert "Number of keyframe nodes in the graph: 20" in r.stdout
    assert "Number of edges in the graph: 3551" in r.stdout


ITER_RE = re.compile(r"^Iter (\d+) // Reprojection error ([-\w.+]+) // Cost ([-\w.+]+) // n relins: (\d+) // n robust edges (\d+)$", re.M)
SLAM_RE = re.compile(r"^Iters (\d+) \(since last kf (\d+)\) // Reprojection error ([-\w.+]+) // Cost ([-\w.+]+) // n relins: (\d+) // n robust edges (\d+)$", re.M)


@pytest.mark.gpu
def test_ba_log_matches_engine(tmp_path):
    path = write_fixture("fr1xyz", tmp_path)
    env = dict(os.environ, GC_PROFILE_LOG_DIR=str(tmp_path))
    r = run([BA, "--bal_file", path, "--n_iters", "80", "--profile", "true"], env=env)
    assert r.returncode == 0, r.stderr
    for line in ("Completed loading data!", "Bundle Adjustment", "Number of keyframe nodes in the graph: 42",
                 "Number of landmark nodes in the graph: 2194", "Number of edges in the graph: 12908",
                 "Sending priors and computing factor potentials.", "Number of iterations: 80", " Finished GBP."):
        assert line in r.stdout, line
    assert r.stdout.count("Weakening priors") == 5
    rows = ITER_RE.findall(r.stdout)
    assert [int(x[0]) for x in rows] == list(range(80))
    st = common.make_setup("fr1xyz")
    gpu = GBPEngine(st.problem)
    init = gpu.eval()
    m = re.search(r"Initial Reprojection error: ([-\w.+]+) Cost ([-\w.+]+)", r.stdout)
    assert float(m.group(1)) == pytest.approx(init["reproj_mean"], rel=1e-5)
    assert float(m.group(1)) == pytest.approx(199.1097, rel=1e-3)        # SURVEY.md 8c known answer
    for it in range(80):
        if (it + 1) % 2 == 0 and it < 10:
            gpu.weaken_priors()
        s = gpu.iterate(1, stats=True)[0]
        _, e, c, nr, nb = rows[it]
        assert float(e) == pytest.approx(s["reproj_mean"], rel=1e-5), it
        assert float(c) == pytest.approx(s["cost"], rel=1e-5), it
        assert (int(nr), int(nb)) == (s["n_relins"], s["n_robust"]), it
    assert os.path.exists(os.path.join(str(tmp_path), "gbp_profile.json"))


@pytest.mark.gpu
def test_slam_log_matches_engine(tmp_path):
    path = write_fixture("fr2robot2", tmp_path)
    r = run([SLAM, "--bal_file", path, "--iters_between_kfs", "25"])
    assert r.returncode == 0, r.stderr
    assert "Loaded data onto host!" in r.stdout and "SLAM" in r.stdout
    assert "Total number of GBP iterations: 474" in r.stdout            # (20-1)*25-1, slam.cpp:1013
    assert "GBP iterations between sucessive keyframes: 25" in r.stdout
    assert [int(x) for x in re.findall(r"Adding keyframe (\d+)", r.stdout)] == list(range(2, 20))
    assert [int(x) for x in re.findall(r"Adding (\d+) new landmarks", r.stdout)] == \
        [31, 19, 33, 17, 29, 19, 22, 19, 25, 40, 54, 34, 59, 102, 73, 41, 19, 0]
    rows = SLAM_RE.findall(r.stdout)
    assert len(rows) == 474
    assert [int(x[1]) for x in rows[:24]] == list(range(24)) and int(rows[24][1]) == 0 and int(rows[24][0]) == 25
    st = common.make_setup("fr2robot2", mode=MODE_SLAM)
    gpu = GBPEngine(st.problem)
    finals = common.slam_run(gpu, st, 25)
    assert float(rows[-1][2]) == pytest.approx(finals[-1]["reproj_mean"], rel=1e-5)
    assert float(rows[23][2]) == pytest.approx(finals[0]["reproj_mean"], rel=1e-5)
    # keyframes are inserted on the device by default; the reference's host round trip gives the same log
    assert "on the device" in r.stdout
    rh = run([SLAM, "--bal_file", path, "--iters_between_kfs", "25", "--host_keyframes", "1"])
    assert rh.returncode == 0, rh.stderr
    assert "READ_PRIORS / host / NEW_KEYFRAME round trip" in rh.stdout
    strip = lambda t: [ln for ln in t.splitlines() if not ln.startswith(("Timing report", "Keyframe insertions"))]
    assert strip(rh.stdout) == strip(r.stdout)


@pytest.mark.gpu
def test_ba_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    path = write_fixture("fr1xyz", tmp_path)
    r1 = run([BA, "--bal_file", path, "--n_iters", "30"])
    r2 = run([BA, "--bal_file", path, "--n_iters", "30", "--ipus", "2"])
    assert r1.returncode == 0 and r2.returncode == 0, r2.stderr
    assert "Number of GPUs: 2" in r2.stdout
    a, b = ITER_RE.findall(r1.stdout), ITER_RE.findall(r2.stdout)
    assert len(a) == len(b) == 30
    for x, y in zip(a[:8], b[:8]):     # same graph; the summation order differs at boundary landmarks (rounding)
        assert float(x[1]) == pytest.approx(float(y[1]), rel=1e-2)
    assert float(b[-1][1]) < 0.5 * float(b[0][1])


@pytest.mark.gpu
def test_ba_converge_and_out_round_trip(tmp_path):
    """--converge stops early on the plateau; --out writes the optimised problem in the input format, and
    loading that file back reproduces the final error as its initial error."""
    path = write_fixture("fr1xyz", tmp_path)
    out = os.path.join(str(tmp_path), "optimised.txt")
    r = run([BA, "--bal_file", path, "--converge", "1e-4", "--out", out])
    assert r.returncode == 0, r.stderr
    rows = ITER_RE.findall(r.stdout)
    m = re.search(r"Stopped after (\d+) iterations: (\w+)", r.stdout)
    assert m and m.group(2) == "converged" and int(m.group(1)) == len(rows) < 1500
    final = float(rows[-1][1])
    assert final < 0.2 * 199.11
    assert f"Optimised problem written to {out}" in r.stdout
    r2 = run([BA, "--bal_file", out, "--n_iters", "0"])
    m2 = re.search(r"Initial Reprojection error: ([-\w.+]+)", r2.stdout)
    assert float(m2.group(1)) == pytest.approx(final, rel=2e-3)
