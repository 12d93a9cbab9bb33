"""Multi-GPU product path (gbp_cuda_init_shard + in-library NCCL exchange), one process per GPU,
against the single-process oracle summing beliefs in the multi-GPU order: bit-identical."""
import numpy as np
import pytest

import common
import oracle_lib
import shard_worker
from test_sharding_cpu import KIND, check_against_global, run_ranks

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("spec,n,exchange", [("seq:fr1xyz", 40, "p2p"), ("synth:64:6000:9:11", 30, "p2p"),
                                              ("synth:64:6000:9:11", 30, "nccl")])
def test_two_gpus_bit_identical_to_oracle_in_sharded_order(tmp_path, spec, n, exchange, monkeypatch):
    """Both exchange paths: peer-to-peer stores over NVLink (CUDA IPC) and the NCCL all-gather."""
    world = 2
    monkeypatch.setenv("GBP_TEST_EXCHANGE", exchange)
    ranks = run_ranks("nccl", world, spec, n, tmp_path)
    st = shard_worker.make_problem(spec)
    ora = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ora.set_shard_bounds(ranks[0]["cam_bounds"])
    stats = []
    for it in range(n):
        common.ba_schedule_step(ora, it)
        stats.append(ora.eval())
    assert int(ranks[0]["n_boundary_points"][0]) > 0
    check_against_global(ranks, ora, st, exact=True)
    # every rank reports the metric of the WHOLE graph, identical on all ranks
    assert np.array_equal(ranks[0]["stats"], ranks[1]["stats"])
    for it in (0, n // 2, n - 1):
        s, o = ranks[0]["stats"][it], stats[it]
        assert s[0] == pytest.approx(o["reproj_mean"], rel=0.01)
        assert (int(s[2]), int(s[3]), int(s[4])) == (o["n_relins"], o["n_robust"], o["n_active"])


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_block_calls_between_processes_bit_identical_to_oracle(tmp_path, world, monkeypatch):
    """One process per GPU, CUDA IPC peer mappings, the call pattern bench.py --gpus N times: blocks of sweeps
    without per-sweep metrics (see tests/test_group_gpu.py for the single-process twin that runs on one GPU)."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    spec, n = "synth:96:9000:9:5", 63
    monkeypatch.setenv("GBP_TEST_EXCHANGE", "p2p")
    monkeypatch.setenv("GBP_TEST_MODE", "blocks")
    ranks = run_ranks("nccl", world, spec, n, tmp_path)
    st = shard_worker.make_problem(spec)
    ora = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ora.set_shard_bounds(ranks[0]["cam_bounds"])
    common.run_ba(ora, 12)
    ora.iterate(n - 12)
    check_against_global(ranks, ora, st, exact=True)


@pytest.mark.skipif(_n_gpus() < 4, reason="needs at least 4 GPUs")
def test_four_gpus_bit_identical_to_oracle_in_sharded_order(tmp_path):
    spec, n = "synth:96:9000:9:5", 24
    ranks = run_ranks("nccl", 4, spec, n, tmp_path)
    st = shard_worker.make_problem(spec)
    ora = oracle_lib.OracleEngine(st.problem, kind=KIND)
    ora.set_shard_bounds(ranks[0]["cam_bounds"])
    common.run_ba(ora, n)
    check_against_global(ranks, ora, st, exact=True)


def test_world_of_one_sharded_handle_equals_plain_handle():
    """The sharded entry point on ONE GPU (gbp_cuda_init_shard with world = 1: the view shard build, the shard-aware
    build, no boundary landmarks) against gbp_cuda_init: every tensor bit-identical after a block of sweeps.  Runs
    in its own process because it creates a NCCL process group."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "scripts", "shard_world1_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    if r.returncode != 0 and "DIFFERENT" not in r.stdout:
        # the process died before comparing anything (seen once on a 4-GPU box right after the 4-rank test, NCCL
        # bootstrap of the one-rank group): show why, try once more; a comparison that ran and differed is never retried
        print("first attempt failed:\n" + (r.stdout + r.stderr)[-3000:])
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "world-1 sharded vs plain: IDENTICAL" in r.stdout, r.stdout[-2000:]
