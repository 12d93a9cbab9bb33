"""gbp_opts.fast_math = 1: the sweep kernel built with contracted multiply-adds.  It is NOT bit-comparable with the
reference (one rounding per a*b+c instead of two, amplified by GBP over the sweeps), so it is held to the north-star
tolerance instead: every message / belief block of ONE sweep from identical state within 1e-4 (block-infinity-norm
relative error) and the plateau reprojection error of a long free run within 1 % of the reference's.
The default (bit-identical) path is what every other GPU test pins."""
import numpy as np
import pytest

import common
import oracle_lib
from gbp_poplar_b200 import GBPEngine, default_opts
from test_sharding_cpu import KIND

pytestmark = pytest.mark.gpu
STATE = ["cam_messages_eta", "cam_messages_lambda", "lmk_messages_eta", "lmk_messages_lambda", "cam_beliefs_eta",
         "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda", "factor_potentials_eta", "factor_potentials_lambda"]


@pytest.mark.parametrize("name,start", [("fr1xyz", 0), ("fr1xyz", 16), ("fr1xyz", 50), ("fr2robot2", 30)])
def test_one_sweep_from_identical_state_within_1e4(name, start):
    """Teacher-forced: the oracle runs `start` sweeps of the ba.cpp schedule, its full state is loaded into a fast-math
    engine, both do ONE sweep."""
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
    worst = 0.0
    for t in STATE:
        d = common.BLOCK_DIMS[t]
        want = common.canon(t, ora.get_tensor(t))
        assert common.canon(t, exact.get_tensor(t)).tobytes() == want.tobytes(), t        # the default path: bit-identical
        err = common.block_rel_err(common.canon(t, fast.get_tensor(t)), want, d)
        worst = max(worst, float(err.max()))
        assert err.max() <= 1e-4, (t, float(err.max()))
    assert worst > 0.0   # the contracted build really is a different rounding (otherwise this test pins nothing)
    fast.close()
    exact.close()


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
