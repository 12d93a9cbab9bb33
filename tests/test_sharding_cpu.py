"""Multi-GPU path, host side: camera-range partition, rank-local sub-problems, boundary lists
(gbp_shard_build) and -- with world_size-2 gloo processes on CPU -- the boundary-landmark
exchange protocol, checked bit for bit against a single-process oracle that sums beliefs in
the multi-GPU order."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import common
import oracle_lib
from gbp_poplar_b200 import BALProblem, Setup
from gbp_poplar_b200.host import Shard

HERE = os.path.dirname(os.path.abspath(__file__))
KIND = "reference" if oracle_lib.available("reference") else "port"


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_ranks(backend, world, spec, n_sweeps, out, timeout=600):
    port = free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "shard_worker.py"), backend, str(r), str(world),
                               str(port), spec, str(n_sweeps), str(out)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = []
    try:
        for p in procs:
            o, _ = p.communicate(timeout=timeout)
            logs.append(o)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{logs[r][-3000:]}"
    return [np.load(os.path.join(out, f"rank{r}.npz")) for r in range(world)]


def check_against_global(ranks, ora, st, exact=True):
    """Every rank's local tensors == the matching slices of the single-process oracle."""
    p = st.problem
    C, L, E = p.n_keyframes, p.n_points, p.n_edges
    cam_ids, lmk_ids = np.array(st.array("cam_ids")), np.array(st.array("lmk_ids"))
    SK, SL = ora.max_nkfedges + 1, ora.max_nlmkedges + 1
    def running_count_global(keys):
        order = np.argsort(keys, kind="stable")
        ks = keys[order]
        first = np.r_[0, np.flatnonzero(np.diff(ks)) + 1]
        start = np.repeat(first, np.diff(np.r_[first, ks.size]))
        out = np.empty(keys.size, np.int64)
        out[order] = np.arange(keys.size) - start
        return out
    slot_c, slot_l = running_count_global(cam_ids.astype(np.int64)), running_count_global(lmk_ids.astype(np.int64))
    G = {t: ora.get_tensor(t) for t in ("cam_beliefs_eta", "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda",
                                        "cam_messages_eta", "cam_messages_lambda", "lmk_messages_eta",
                                        "lmk_messages_lambda", "factor_potentials_eta", "factor_potentials_lambda",
                                        "damping", "damping_count", "robust_flag")}
    seen_edges = np.zeros(E, bool)
    for z in ranks:
        c0, c1 = (int(x) for x in z["cam_range"])
        lg, eg = z["lmk_global"].astype(np.int64), z["edge_global"].astype(np.int64)
        assert not seen_edges[eg].any()
        seen_edges[eg] = True

        def same(a, b, what):
            if exact:
                assert a.tobytes() == b.tobytes(), what
            else:
                assert np.allclose(a, b, rtol=1e-4, atol=1e-6), what
        same(z["cam_beliefs_eta"], G["cam_beliefs_eta"].reshape(C, 6)[c0:c1].ravel(), "cam_beliefs_eta")
        same(z["cam_beliefs_lambda"], G["cam_beliefs_lambda"].reshape(C, 36)[c0:c1].ravel(), "cam_beliefs_lambda")
        same(z["lmk_beliefs_eta"], G["lmk_beliefs_eta"].reshape(L, 3)[lg].ravel(), "lmk_beliefs_eta")
        same(z["lmk_beliefs_lambda"], G["lmk_beliefs_lambda"].reshape(L, 9)[lg].ravel(), "lmk_beliefs_lambda")
        for t, d in (("factor_potentials_eta", 9), ("factor_potentials_lambda", 81), ("damping", 1),
                     ("damping_count", 1), ("robust_flag", 1)):
            same(z[t], G[t].reshape(E, d)[eg].ravel(), t)
        # messages: local slot numbering follows the local edge order = global order restricted to the rank
        lSK = z["cam_messages_eta"].size // (6 * (c1 - c0)) if c1 > c0 else 1
        lSL = z["lmk_messages_eta"].size // (3 * lg.size) if lg.size else 1
        lmk_local = np.full(L, -1, np.int64)
        lmk_local[lg] = np.arange(lg.size)
        ci, li = cam_ids[eg].astype(np.int64) - c0, lmk_local[lmk_ids[eg]]

        def running_count(keys):   # number of earlier local edges with the same key (edges are in order)
            order = np.argsort(keys, kind="stable")
            ks = keys[order]
            first = np.r_[0, np.flatnonzero(np.diff(ks)) + 1]
            start = np.repeat(first, np.diff(np.r_[first, ks.size]))
            out = np.empty(keys.size, np.int64)
            out[order] = np.arange(keys.size) - start
            return out
        lc, ll = running_count(ci), running_count(li)
        for t, d, S, lS, v, s_loc, gv, s_glob in (("cam_messages_eta", 6, SK, lSK, ci, lc, cam_ids[eg], slot_c[eg]),
                                                   ("cam_messages_lambda", 36, SK, lSK, ci, lc, cam_ids[eg], slot_c[eg]),
                                                   ("lmk_messages_eta", 3, SL, lSL, li, ll, lmk_ids[eg], slot_l[eg]),
                                                   ("lmk_messages_lambda", 9, SL, lSL, li, ll, lmk_ids[eg], slot_l[eg])):
            a = common.canon(t, z[t].reshape(-1, lS, d)[v, s_loc + 1].ravel())
            b = common.canon(t, G[t].reshape(-1, S, d)[gv, s_glob + 1].ravel())
            same(a, b, t)
    assert seen_edges.all()


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_shard_build_partitions_the_graph(world):
    st = common.make_setup("fr1xyz")
    p = st.problem
    C, L, E = p.n_keyframes, p.n_points, p.n_edges
    cam_ids, lmk_ids = np.array(st.array("cam_ids")), np.array(st.array("lmk_ids"))
    shards = [Shard(p, world, r, owner=st) for r in range(world)]
    bounds = np.array(shards[0].cam_bounds)
    assert bounds[0] == 0 and bounds[-1] == C and np.all(np.diff(bounds.astype(np.int64)) >= 0)
    edges = np.concatenate([np.array(s.edge_global) for s in shards])
    assert np.array_equal(np.sort(edges), np.arange(E))                      # every factor on exactly one rank
    rank_of_cam = np.searchsorted(bounds[1:], np.arange(C), side="right")
    n_ranks_of_lmk = np.array([len(set(rank_of_cam[cam_ids[lmk_ids == l]])) for l in range(L)])
    boundary_global = np.flatnonzero(n_ranks_of_lmk > 1)
    loads = []
    for r, s in enumerate(shards):
        q = s.problem
        eg, lg = np.array(s.edge_global), np.array(s.lmk_global)
        assert (q.n_keyframes, q.n_points, q.n_edges) == (bounds[r + 1] - bounds[r], lg.size, eg.size)
        assert s.n_boundary_points == boundary_global.size
        lc = np.ctypeslib.as_array(q.cam_ids, shape=(eg.size,)) if eg.size else np.zeros(0, np.uint32)
        ll = np.ctypeslib.as_array(q.lmk_ids, shape=(eg.size,)) if eg.size else np.zeros(0, np.uint32)
        assert np.array_equal(lc + bounds[r], cam_ids[eg]) and np.array_equal(lg[ll], lmk_ids[eg])
        assert np.all(np.diff(eg.astype(np.int64)) > 0) and np.all(np.diff(lg.astype(np.int64)) > 0)
        z = np.ctypeslib.as_array(q.measurements, shape=(2 * eg.size,)).reshape(-1, 2)
        assert np.array_equal(z, np.array(st.array("measurements")).reshape(-1, 2)[eg])
        pe = np.ctypeslib.as_array(q.lmk_priors_eta, shape=(3 * lg.size,)).reshape(-1, 3)
        assert np.array_equal(pe, np.array(st.array("lmk_priors_eta")).reshape(-1, 3)[lg])
        # boundary lists: the touched subset of the global boundary list, with its positions
        bl, bs = np.array(s.boundary_local), np.array(s.boundary_slot)
        assert np.array_equal(boundary_global[bs], lg[bl])
        assert set(lg[bl]) == set(boundary_global) & set(lg)
        assert s.n_active_global == E
        # the ranks observing every boundary landmark (who sends a partial sum to whom)
        masks = np.array(s.boundary_ranks)
        for g_l, m in zip(lg[bl], masks):
            ranks = set(int(x) for x in rank_of_cam[cam_ids[lmk_ids == g_l]])
            assert m == sum(1 << x for x in ranks) and (m >> r) & 1 and len(ranks) > 1
        loads.append(eg.size)
        plan = _plan(p, world, r)
        assert (plan.cam_begin, plan.cam_end, plan.n_local_edges, plan.n_local_points, plan.n_boundary_points) == \
               (bounds[r], bounds[r + 1], eg.size, lg.size, boundary_global.size)
    if world > 1:
        assert max(loads) < 1.35 * E / world   # balanced by warp-tile count (camera granularity)


_FIELDS = [  # (member of gbp_problem, elements per local edge / camera / landmark, which count)
    ("cam_ids", 1, "E"), ("lmk_ids", 1, "E"), ("measurements", 2, "E"), ("meas_variances", 1, "E"),
    ("active_flag", 1, "E"), ("damping", 1, "E"), ("damping_count", 1, "E"), ("mu", 9, "E"), ("oldmu", 9, "E"),
    ("cam_priors_eta", 6, "C"), ("cam_priors_lambda", 36, "C"), ("cam_scaling", 1, "C"), ("cam_weaken_flag", 1, "C"),
    ("lmk_priors_eta", 3, "L"), ("lmk_priors_lambda", 9, "L"), ("lmk_scaling", 1, "L"), ("lmk_weaken_flag", 1, "L"),
]


def _problem_arrays(q):
    n = {"E": q.n_edges, "C": q.n_keyframes, "L": q.n_points}
    out = {}
    for name, width, kind in _FIELDS:
        ptr = getattr(q, name)
        out[name] = None if not ptr or n[kind] == 0 else np.ctypeslib.as_array(ptr, shape=(width * n[kind],)).copy()
    return out


@pytest.mark.parametrize("sorted_by_camera", [True, False])
@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_shard_view_equals_owning_shard(world, sorted_by_camera):
    """gbp_shard_build_view (what gbp_cuda_init_shard uses: contiguous runs of the global arrays are referenced, not
    copied) describes exactly the same sub-problem as gbp_shard_build, also when the edge list is not camera-sorted
    (nothing is contiguous then and everything is copied) and when mu / oldmu are not all zero."""
    st = common.make_setup("fr2robot2")
    p = st.problem
    E = p.n_edges
    keep = [st]
    if not sorted_by_camera:
        bal = common.load_sequence("fr2robot2")
        perm = np.random.default_rng(3).permutation(bal.n_edges)
        params = bal.parameters
        C = bal.n_keyframes
        shuffled = BALProblem.from_arrays(bal.intrinsics, bal.camera_index[perm], bal.point_index[perm],
                                          bal.observations.reshape(-1, 2)[perm], params[:6 * C], params[6 * C:])
        st = Setup(shuffled)
        p = st.problem
        keep += [shuffled, st]
    # the setup leaves mu / oldmu NULL (= the zeros the reference streams): give a copy of the problem a non-zero mu
    import ctypes as C
    from gbp_poplar_b200 import _capi
    p = _capi.GbpProblem.from_buffer_copy(p)
    mu = np.random.default_rng(1).normal(size=9 * E).astype(np.float32)
    oldmu = np.zeros(9 * E, np.float32)                                         # oldmu stays all zero
    p.mu, p.oldmu = mu.ctypes.data_as(_capi.c_f32p), oldmu.ctypes.data_as(_capi.c_f32p)
    keep += [mu, oldmu]
    for r in range(world):
        own, view = Shard(p, world, r, owner=st), Shard(p, world, r, owner=st, view=True)
        a, b = _problem_arrays(own.problem), _problem_arrays(view.problem)
        assert (own.problem.n_keyframes, own.problem.n_points, own.problem.n_edges) == \
               (view.problem.n_keyframes, view.problem.n_points, view.problem.n_edges)
        for name, _, _ in _FIELDS:
            assert (a[name] is None) == (b[name] is None), name
            if a[name] is not None:
                assert a[name].tobytes() == b[name].tobytes(), name
        assert a["oldmu"] is None                                      # all zero: left to the library's default
        if own.problem.n_edges:
            assert a["mu"] is not None and np.array_equal(a["mu"].reshape(-1, 9), mu.reshape(-1, 9)[np.array(own.edge_global)])
        for f in ("edge_global", "lmk_global", "boundary_local", "boundary_slot", "cam_bounds"):
            assert np.array_equal(np.array(getattr(own, f)), np.array(getattr(view, f))), f
        assert (own.n_boundary_points, own.n_active_global) == (view.n_boundary_points, view.n_active_global)


def _plan(p, world, rank):
    import ctypes as C
    from gbp_poplar_b200 import _capi
    out = _capi.GbpShardPlan()
    assert _capi.load_library().gbp_cuda_plan_shard(C.byref(p), world, rank, C.byref(out)) == 0
    return out


def test_shard_rejects_bad_arguments():
    import ctypes as C
    from gbp_poplar_b200 import _capi
    st = common.make_setup("fr2robot2")
    lib = _capi.load_library()
    h = C.c_void_p()
    assert lib.gbp_shard_build(C.byref(st.problem), 0, 0, C.byref(h)) != 0
    assert lib.gbp_shard_build(C.byref(st.problem), 2, 2, C.byref(h)) != 0


def test_oracle_sharded_order_equals_tile_order_for_one_rank():
    st = common.make_setup("fr2robot2")
    a = oracle_lib.OracleEngine(st.problem, kind=KIND)
    a.set_reduce_order(1)
    b = oracle_lib.OracleEngine(st.problem, kind=KIND)
    b.set_shard_bounds([0, st.problem.n_keyframes])
    common.run_ba(a, 25)
    common.run_ba(b, 25)
    for t in ("cam_beliefs_lambda", "lmk_beliefs_lambda", "lmk_messages_eta"):
        assert a.get_tensor(t).tobytes() == b.get_tensor(t).tobytes(), t


@pytest.mark.parametrize("spec,world,n", [("seq:fr2robot2", 2, 30), ("synth:24:1500:8:7", 2, 24),
                                          ("synth:24:1500:8:7", 3, 12)])
def test_gloo_ranks_match_single_process_oracle_bit_for_bit(tmp_path, spec, world, n):
    """world_size-N gloo run of the exchange protocol == single-process oracle in the multi-GPU order."""
    import shard_worker
    ranks = run_ranks("gloo", world, spec, n, tmp_path)
    st = shard_worker.make_problem(spec)
    ora = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ora.set_shard_bounds(ranks[0]["cam_bounds"])
    common.run_ba(ora, n)
    assert int(ranks[0]["n_boundary_points"][0]) > 0
    check_against_global(ranks, ora, st, exact=True)
    # and the partition only perturbs rounding: close to the serial single-GPU order after a few sweeps
    ser = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ser.set_reduce_order(1)
    common.run_ba(ser, 3)
    ora3 = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ora3.set_shard_bounds(ranks[0]["cam_bounds"])
    common.run_ba(ora3, 3)
    err = common.block_rel_err(ora3.get_tensor("lmk_beliefs_lambda"), ser.get_tensor("lmk_beliefs_lambda"), 9)
    assert err.max() < 1e-4
