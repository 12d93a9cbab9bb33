"""Helpers shared by the tests: fixtures loading, schedules, comparisons."""
import os

import numpy as np

from gbp_poplar_b200 import BALProblem, Setup, cli_options, MODE_BA, MODE_SLAM

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

BLOCK_DIMS = {
    "cam_beliefs_eta": 6, "cam_beliefs_lambda": 36, "lmk_beliefs_eta": 3, "lmk_beliefs_lambda": 9,
    "cam_messages_eta": 6, "cam_messages_lambda": 36, "lmk_messages_eta": 3, "lmk_messages_lambda": 9,
    "factor_potentials_eta": 9, "factor_potentials_lambda": 81,
}


# Tensors of which the CUDA path stores only what the algorithm reads (gbp_layout.h): the strict
# upper triangle of a factor->camera message Lambda is never read back (inv6x6 uses the lower
# triangle, matlib.cpp:193-206; the belief sum is formed on chip from the full message), so
# get_tensor mirrors the lower triangle.  canon() maps both sides to that form.
LOWER_ONLY = ("cam_messages_lambda", "pcam_messages_lambda")


def canon(name, a):
    if name not in LOWER_ONLY:
        return a
    m = np.array(a).reshape(-1, 6, 6)
    il = np.tril_indices(6, -1)
    m[:, il[1], il[0]] = m[:, il[0], il[1]]
    return m.reshape(-1)


def load_sequence(name):
    """BALProblem of one of the reference sequences, from the frozen input fixture."""
    z = np.load(os.path.join(GOLDEN, f"seq_{name}.npz"))
    C, L, E = (int(x) for x in z["dims"])
    params = z["parameters"]
    return BALProblem.from_arrays(z["intrinsics"], z["cam_idx"].astype(np.uint32), z["lmk_idx"].astype(np.uint32),
                                  z["observations"], params[:6 * C], params[6 * C:])


def make_setup(name, mode=MODE_BA, **opts):
    bal = load_sequence(name)
    return Setup(bal, cli_options(**opts), mode)


def ba_schedule_step(engine, it, steps=5):
    """One iteration of the ba.cpp loop body (ba/ba.cpp:1003-1008)."""
    if (it + 1) % 2 == 0 and it < steps * 2:
        engine.weaken_priors()
    engine.iterate(1)


def run_ba(engine, n_iters, steps=5, start=0):
    for it in range(start, start + n_iters):
        ba_schedule_step(engine, it, steps)


def block_rel_err(a, b, d):
    """max over blocks of  max|a-b| / max|b|  (block-infinity-norm relative error)."""
    a = np.asarray(a, np.float64).reshape(-1, d)
    b = np.asarray(b, np.float64).reshape(-1, d)
    num = np.abs(a - b).max(axis=1)
    den = np.abs(b).max(axis=1)
    ok = den > 0
    out = np.zeros_like(num)
    out[ok] = num[ok] / den[ok]
    out[~ok] = num[~ok]
    return out


def slam_run(engine, setup, iters_between_kfs, steps=5, stats_every_kf=True, on_kf=None, device_kf=False):
    """The slam.cpp loop (ba/slam.cpp:1013-1103) against any engine.  device_kf: keyframe insertion
    through gbp_cuda_add_keyframe_device instead of the READ_PRIORS / host / NEW_KEYFRAME round trip."""
    C = setup.problem.n_keyframes
    data_counter = 0
    niters = (C - 1) * iters_between_kfs - 1
    it = 0
    finals = []
    for i in range(niters):
        if (i + 1) % iters_between_kfs == 0:
            if stats_every_kf:
                finals.append(engine.eval())
            it = 0
            data_counter += 1
            if device_kf:
                n_new = engine.add_keyframe_device(data_counter + 1, steps)
                if on_kf:
                    on_kf(data_counter, n_new)
                ba_schedule_step(engine, it, steps)
                it += 1
                continue
            b = engine.get_beliefs()
            pr = engine.get_priors()
            n_new, dc = setup.next_keyframe(b["cam_beliefs_eta"], b["cam_beliefs_lambda"], pr["cam_priors_eta"],
                                            pr["cam_priors_lambda"], pr["lmk_priors_eta"], pr["lmk_priors_lambda"])
            engine.add_keyframe(dc, pr["cam_priors_eta"], pr["cam_priors_lambda"], pr["lmk_priors_eta"],
                                pr["lmk_priors_lambda"], setup.array("active_flag"), setup.array("cam_weaken_flag"),
                                setup.array("lmk_weaken_flag"))
            if on_kf:
                on_kf(setup.data_counter, n_new)
        ba_schedule_step(engine, it, steps)
        it += 1
    finals.append(engine.eval())
    return finals
