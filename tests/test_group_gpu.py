"""The multi-GPU product path on the handles the benchmark times -- CUDA-graph replay of block calls without
per-sweep metrics (upper-triangle skipping, device-side exchange step counter), the fused peer-to-peer boundary
exchange, the metric exchange over peer memory -- driven as a single-process group (gbp_cuda_init_group).  The
shards SHARE device 0, so these tests run on a one-GPU box; with more GPUs visible the same tests also run one
shard per device.  Checked bit for bit against the single-process oracle summing beliefs in the multi-GPU
order."""
import numpy as np
import pytest

import common
import oracle_lib
import shard_worker
from conftest import cuda_device_count
from test_sharding_cpu import KIND, check_against_global

pytestmark = pytest.mark.gpu


def dump_rank(eng):
    """What shard_worker writes for one rank: its local tensors + index maps."""
    res = {t: eng.get_tensor(t) for t in shard_worker.DUMP}
    sh = eng.shard
    res["lmk_global"] = np.array(sh.lmk_global)
    res["edge_global"] = np.array(sh.edge_global)
    res["cam_range"] = np.array([sh.cam_begin, sh.cam_end])
    res["cam_bounds"] = np.array(sh.cam_bounds)
    res["boundary_local"] = np.array(sh.boundary_local)
    res["n_boundary_points"] = np.array([sh.n_boundary_points])
    return res


def device_layouts(world):
    """All shards on device 0; one shard per device as well when the box has enough GPUs."""
    out = [[0] * world]
    if cuda_device_count() >= world:
        out.append(list(range(world)))
    return out


def make_group(st, world, devices, **opts):
    from gbp_poplar_b200 import GBPGroup, default_opts
    grp = GBPGroup(st.problem, world, devices=devices, opts=default_opts(**opts))
    assert all(r.exchange_mode() == "p2p" for r in grp.ranks)
    return grp


@pytest.mark.parametrize("spec,world,blocks", [
    ("synth:64:6000:9:11", 2, (7, 1, 13, 30)),     # 3490 boundary landmarks; sweeps 12..62 cross several lock-step relinearisations
    ("seq:fr1xyz", 3, (7, 1, 13, 30)),
    ("seq:fr2robot2", 4, (5, 2, 17)),
])
@pytest.mark.parametrize("protocol", ["pull", "push"])
def test_block_calls_without_stats_bit_identical_to_oracle(spec, world, blocks, protocol, monkeypatch):
    """The path bench.py --gpus N times: gbp_cuda_iterate(n) blocks with stats == NULL on sharded handles.  Both
    directions of the boundary exchange: partial sums read from the peers' buffers (the default) and stored into
    them (GBP_XCHG_PUSH=1, read when a handle is built)."""
    monkeypatch.setenv("GBP_XCHG_PUSH", "1" if protocol == "push" else "0")
    st = shard_worker.make_problem(spec)
    for devices in device_layouts(world):
        grp = make_group(st, world, devices)
        ora = oracle_lib.OracleEngine(st.problem, kind=KIND)
        ora.set_shard_bounds(np.array(grp.ranks[0].shard.cam_bounds))
        assert grp.ranks[0].shard.n_boundary_points > 0
        common.run_ba(grp, 12)          # the ba.cpp schedule incl. prior weakening, one sweep per call
        common.run_ba(ora, 12)
        check_against_global([dump_rank(r) for r in grp.ranks], ora, st, exact=True)
        for n in blocks:
            grp.iterate(n)              # graph replay: n - 1 lower-only sweeps + one complete sweep
            ora.iterate(n)
            check_against_global([dump_rank(r) for r in grp.ranks], ora, st, exact=True)
        ms, launches = grp.last_timing()
        assert ms > 0 and launches == 2 * world * blocks[-1], (ms, launches)   # k_sweep + k_update_vars per rank and sweep
        grp.close()


@pytest.mark.parametrize("relin_mode", [1, 2])
def test_both_sweep_flavours_on_shards(relin_mode):
    st = shard_worker.make_problem("synth:64:6000:9:11")
    grp = make_group(st, 2, [0, 0], relin_mode=relin_mode)
    ora = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ora.set_shard_bounds(np.array(grp.ranks[0].shard.cam_bounds))
    common.run_ba(grp, 12)
    common.run_ba(ora, 12)
    for n in (9, 16):
        grp.iterate(n)
        ora.iterate(n)
    check_against_global([dump_rank(r) for r in grp.ranks], ora, st, exact=True)
    grp.close()


def test_per_sweep_metric_over_peer_memory():
    """stats != NULL on shards: complete sweeps + the five metric sums exchanged through the peers' receive
    buffers (no NCCL), replayed from a CUDA graph; every rank reports the metric of the WHOLE graph."""
    st = shard_worker.make_problem("seq:fr1xyz")
    grp = make_group(st, 2, [0, 0])
    ora = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ora.set_shard_bounds(np.array(grp.ranks[0].shard.cam_bounds))
    got, want = [], []
    for it in range(30):
        if (it + 1) % 2 == 0 and it < 10:
            grp.weaken_priors()
            ora.weaken_priors()
        got.append(grp.iterate(1, stats=True)[0])
        ora.iterate(1)
        want.append(ora.eval())
    got += grp.iterate(11, stats=True)          # a block of sweeps with per-sweep metrics
    for _ in range(11):
        ora.iterate(1)
        want.append(ora.eval())
    check_against_global([dump_rank(r) for r in grp.ranks], ora, st, exact=True)
    for g, w in zip(got, want):
        assert g["reproj_mean"] == pytest.approx(w["reproj_mean"], rel=1e-3)   # the device metric solves in double
        assert (g["n_relins"], g["n_robust"], g["n_active"]) == (w["n_relins"], w["n_robust"], w["n_active"])
    assert grp.eval()["reproj_mean"] == pytest.approx(want[-1]["reproj_mean"], rel=1e-3)
    grp.close()


def test_group_of_one_is_a_plain_handle():
    from gbp_poplar_b200 import GBPEngine, GBPGroup
    st = shard_worker.make_problem("seq:fr2robot2")
    grp = GBPGroup(st.problem, 1, devices=[0])      # a group of one never exchanges
    eng = GBPEngine(st.problem)
    common.run_ba(grp, 12)
    common.run_ba(eng, 12)
    grp.iterate(9)
    eng.iterate(9)
    for t in shard_worker.DUMP:
        assert grp.ranks[0].get_tensor(t).tobytes() == eng.get_tensor(t).tobytes(), t
    grp.close()
    eng.close()


def test_shards_sharing_a_device_refuse_a_boundary_that_could_starve_the_peer():
    from gbp_poplar_b200 import GBPGroup
    st = shard_worker.make_problem("synth:200:20000:10:3")   # 3 ranks: 7722 boundary landmarks on the middle rank
    with pytest.raises(RuntimeError, match="too many boundary landmarks"):
        GBPGroup(st.problem, 3, devices=[0, 0, 0])


def test_problem_arrays_may_be_freed_after_init():
    """gbp_cuda_init_group / _init_shard only read the caller's arrays during the call: the retained shard keeps
    no pointer into them (its view members are NULL)."""
    st = shard_worker.make_problem("seq:fr2robot2")
    grp = make_group(st, 2, [0, 0])
    for r in grp.ranks:
        q = r.shard.problem
        assert not q.measurements and not q.cam_priors_eta and not q.meas_variances
        assert q.n_edges == r.n_edges and bool(q.cam_ids) and bool(q.lmk_ids)   # owned index maps stay
    grp.close()
