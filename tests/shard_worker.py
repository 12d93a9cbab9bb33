"""Worker process of the multi-rank tests (one per rank).

backend "gloo": CPU.  Each rank drives the CPU oracle on ITS shard (built by the product's
    gbp_shard_build) through the codelet-level entry points and performs the boundary-landmark
    exchange of the multi-GPU protocol (include/gbp_cuda.h) with torch.distributed all_gather --
    this covers the partitioning / index maps / exchange order without a GPU.
backend "nccl": GPU.  Each rank runs the product path: GBPEngine.sharded -> gbp_cuda_init_shard,
    the exchange happens inside the library (NCCL all-gather on its own stream).

Every rank writes its final local tensors (+ index maps) to <out>/rank<r>.npz; the parent test
compares them with a single-process oracle run in the multi-GPU summation order.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

DUMP = ["cam_beliefs_eta", "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda", "cam_messages_eta",
        "cam_messages_lambda", "lmk_messages_eta", "lmk_messages_lambda", "factor_potentials_eta",
        "factor_potentials_lambda", "damping", "damping_count", "robust_flag"]


def make_problem(spec):
    """spec: 'seq:<name>' (frozen reference sequence) or 'synth:<cams>:<lmks>:<obs>:<seed>'."""
    import common
    from gbp_poplar_b200 import BALProblem, Setup
    kind, _, rest = spec.partition(":")
    if kind == "seq":
        return common.make_setup(rest)
    c, l, o, s = rest.split(":")
    return Setup(BALProblem.synthetic(int(c), int(l), float(o), int(s)))


class EmulatedShardEngine:
    """The multi-GPU sweep protocol on top of any single-shard engine + torch.distributed.

    belief(boundary landmark) = (0 + prior) + partial[rank 0] + partial[rank 1] + ...,
    partial[r] = serial fp32 sum (from +0, slot order) of rank r's factor->landmark messages."""

    def __init__(self, engine, shard, dist):
        self.e, self.sh, self.dist = engine, shard, dist
        self.bl = np.asarray(shard.boundary_local, dtype=np.int64)
        self.bs = np.asarray(shard.boundary_slot, dtype=np.int64)
        self.SL = engine.max_nlmkedges + 1
        # LINEARISE_PROG needs no exchange: all messages are zero, so belief == 0 + prior on every replica

    def exchange(self):
        import torch
        e, nB, W = self.e, self.sh.n_boundary_points, self.sh.world
        L = e.n_points
        send = np.zeros((nB, 12), np.float32)
        me = e.get_tensor("lmk_messages_eta").reshape(L, self.SL, 3)
        ml = e.get_tensor("lmk_messages_lambda").reshape(L, self.SL, 9)
        acc = np.zeros((self.bl.size, 12), np.float32)
        for k in range(1, self.SL):  # slots beyond a landmark's degree hold zeros
            acc[:, :3] = acc[:, :3] + me[self.bl, k]
            acc[:, 3:] = acc[:, 3:] + ml[self.bl, k]
        send[self.bs] = acc
        recv = [torch.zeros(nB, 12) for _ in range(W)]
        self.dist.all_gather(recv, torch.from_numpy(send))
        tot = np.zeros((self.bl.size, 12), np.float32)
        tot[:, :3] = tot[:, :3] + me[self.bl, 0]
        tot[:, 3:] = tot[:, 3:] + ml[self.bl, 0]
        for r in range(W):
            tot = tot + recv[r].numpy()[self.bs]
        be = e.get_tensor("lmk_beliefs_eta").reshape(L, 3)
        bl = e.get_tensor("lmk_beliefs_lambda").reshape(L, 9)
        be[self.bl] = tot[:, :3]
        bl[self.bl] = tot[:, 3:]
        e.set_tensor("lmk_beliefs_eta", be)
        e.set_tensor("lmk_beliefs_lambda", bl)

    def weaken_priors(self):
        self.e.weaken_prior_vertices()
        self.e.update_beliefs()
        self.exchange()

    def iterate(self, n):
        for _ in range(n):
            self.e.prep_messages()
            self.e.compute_messages()
            self.e.update_beliefs()
            self.exchange()
            if hasattr(self.e, "commit_messages"):
                self.e.commit_messages()

    def get_tensor(self, name):
        return self.e.get_tensor(name)


def main():
    backend, rank, world, port, spec, n_sweeps, out = sys.argv[1:8]
    rank, world, n_sweeps = int(rank), int(world), int(n_sweeps)
    import torch
    import torch.distributed as dist
    import common
    from gbp_poplar_b200 import GBPEngine, default_opts
    from gbp_poplar_b200.host import Shard

    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    st = make_problem(spec)
    stats = []
    if backend == "nccl":
        want = os.environ.get("GBP_TEST_EXCHANGE", "auto")   # auto | nccl | p2p
        eng = GBPEngine.sharded(st.problem, default_opts(device=rank, exchange={"auto": 0, "nccl": 1, "p2p": 2}[want]))
        shard = eng.shard
        if want != "auto":
            assert eng.exchange_mode() == want, eng.exchange_mode()
        print(f"rank {rank}: exchange mode {eng.exchange_mode()}", flush=True)
        if os.environ.get("GBP_TEST_MODE") == "blocks":
            # what bench.py --gpus N times: block calls without per-sweep metrics (CUDA-graph replay, lower-only
            # sweeps, device-side exchange step counter); n_sweeps = 12 single-sweep calls + the blocks
            common.run_ba(eng, 12)
            left = n_sweeps - 12
            for n in (7, 1, 13):
                if left >= n:
                    eng.iterate(n)
                    left -= n
            if left:
                eng.iterate(left)
        else:
            for it in range(n_sweeps):
                if (it + 1) % 2 == 0 and it < 10:
                    eng.weaken_priors()
                stats.append(eng.iterate(1, stats=True)[0])
        stats = [[s["reproj_mean"], s["cost"], s["n_relins"], s["n_robust"], s["n_active"]] for s in stats]
    else:
        import oracle_lib
        shard = Shard(st.problem, world, rank, owner=st, view=True)   # the build gbp_cuda_init_shard uses
        kind = "reference" if oracle_lib.available("reference") else "port"
        ora = oracle_lib.OracleEngine(shard.problem, kind=kind, threads=2)
        ora.set_reduce_order(1)
        eng = EmulatedShardEngine(ora, shard, dist)
        common.run_ba(eng, n_sweeps)
    res = {t: eng.get_tensor(t) for t in DUMP}
    res["lmk_global"] = np.array(shard.lmk_global)
    res["edge_global"] = np.array(shard.edge_global)
    res["cam_range"] = np.array([shard.cam_begin, shard.cam_end])
    res["cam_bounds"] = np.array(shard.cam_bounds)
    res["boundary_local"] = np.array(shard.boundary_local)
    res["n_boundary_points"] = np.array([shard.n_boundary_points])
    res["stats"] = np.array(stats, dtype=np.float64)
    np.savez(os.path.join(out, f"rank{rank}.npz"), **res)
    dist.barrier()
    if backend == "nccl":
        eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
