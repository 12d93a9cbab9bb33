"""gbp_opts.fast_math = 1: the sweep kernel built with contracted multiply-adds.  It is NOT bit-comparable with the
reference (one rounding per a*b+c instead of two), and what it can be held to was MEASURED (scripts/fast_math_errors.py,
profiles/r2_fast_math_one_sweep_errors.log), block-infinity-norm relative error after ONE sweep from the reference's state:

  * from the initial state every message / belief block agrees within 1e-4 (max 1.4e-5): the north-star tolerance holds;
  * from mid-run states (sweep 16 / 50 / 300) it holds for the camera beliefs (<= 6e-6) and for >= 97 % of the message
    blocks, but the landmark-bound eta -- a small difference of large terms in the Schur complement -- amplifies the
    single rounding difference to 1e-3 .. 3e-1 on 0.04 .. 3 % of the blocks.  That is why the DEFAULT is the
    order-faithful, never-contracted build (every other GPU test pins it bit for bit) and this one is opt-in;
  * the plateau reprojection error of the reference's 1500-sweep fr1xyz run stays within 1 %.
"""
import numpy as np
import pytest

import common
import oracle_lib
from gbp_poplar_b200 import GBPEngine, default_opts
from test_sharding_cpu import KIND

pytestmark = pytest.mark.gpu
STATE = ["cam_messages_eta", "cam_messages_lambda", "lmk_messages_eta", "lmk_messages_lambda", "cam_beliefs_eta",
         "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda", "factor_potentials_eta", "factor_potentials_lambda"]


def _one_sweep_errors(name, start):
    """Teacher-forced: the oracle runs `start` sweeps of the ba.cpp schedule, its full state is loaded into a fast-math and
    a default engine, all three do ONE sweep.  Returns {tensor: block errors of the fast-math engine}."""
    st = common.make_setup(name)
    ora = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ora.set_reduce_order(1)
    common.run_ba(ora, start)
    fast = GBPEngine(st.problem, default_opts(fast_math=1))
    exact = GBPEngine(st.problem)
    snap = ora.snapshot()
    fast.restore(snap)
    exact.restore(snap)
    ora.iterate(1)
    fast.iterate(1)
    exact.iterate(1)
    out = {}
    for t in STATE:
        want = common.canon(t, ora.get_tensor(t))
        assert common.canon(t, exact.get_tensor(t)).tobytes() == want.tobytes(), t        # the default path: bit-identical
        out[t] = common.block_rel_err(common.canon(t, fast.get_tensor(t)), want, common.BLOCK_DIMS[t])
    fast.close()
    exact.close()
    return out


def test_one_sweep_from_the_initial_state_within_1e4():
    errs = _one_sweep_errors("fr1xyz", 0)
    for t, err in errs.items():
        assert err.max() <= 1e-4, (t, float(err.max()))
    assert max(float(e.max()) for e in errs.values()) > 0.0   # it really is a different rounding


@pytest.mark.parametrize("name,start", [("fr1xyz", 16), ("fr1xyz", 50), ("fr2robot2", 30)])
def test_one_sweep_from_mid_run_states(name, start):
    """What holds mid-run (see the module docstring): camera beliefs tight, the bulk of every tensor within 1e-4, the
    ill-conditioned tail of the landmark-bound eta bounded."""
    errs = _one_sweep_errors(name, start)
    for t in ("cam_beliefs_eta", "cam_beliefs_lambda"):
        assert errs[t].max() <= 1e-3, (t, float(errs[t].max()))
    for t, err in errs.items():
        assert np.median(err) <= 1e-4, (t, float(np.median(err)))
        if t not in ("lmk_beliefs_eta", "lmk_messages_eta"):
            assert np.percentile(err, 99) <= 1e-3, (t, float(np.percentile(err, 99)))
        assert err.max() <= 0.5, (t, float(err.max()))


def test_plateau_reprojection_error_within_one_percent():
    """fr1xyz, the reference's default 1500-sweep BA run: error on the plateau (sweeps 499 / 999 / 1499)."""
    st = common.make_setup("fr1xyz")
    fast = GBPEngine(st.problem, default_opts(fast_math=1))
    exact = GBPEngine(st.problem)
    got, want = {}, {}
    done = 0
    for stop in (500, 1000, 1500):
        for eng, out in ((fast, got), (exact, want)):
            it = done
            while it < stop:
                if it < 10:
                    common.ba_schedule_step(eng, it)
                    it += 1
                else:
                    eng.iterate(stop - it)
                    it = stop
            out[stop] = eng.eval()["reproj_mean"]
        done = stop
    for k in (500, 1000, 1500):
        assert got[k] == pytest.approx(want[k], rel=0.01), (k, got, want)
    fast.close()
    exact.close()
