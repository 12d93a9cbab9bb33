#!/usr/bin/env python
"""bench.py -- GBP sweeps/s and factor-message updates/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU codelets

A "step" is ONE GBP sweep (GBP_PROG, ba/ba.cpp:895-905: prep -> messages -> belief
update) over the whole factor graph.  Workload at N=1 = BASELINE.json configs[3]:
synthetic BAL-format problem, 1k cameras / 100k landmarks / ~1M reprojection factors
(generated in-process, seed 1234; no dataset is read).  At N>1 the SAME generator makes
ONE graph N times as large (N k cameras / N x 100k landmarks / ~N M factors -- N=8 is the
scale of configs[4]) which is partitioned by camera range over the N GPUs, one process
per GPU, with the per-sweep boundary-landmark exchange over NVLink (weak scaling).  `value` = factor-message
updates per second (= factors x sweeps / s, one update = both directed messages of one
factor), device-timed with CUDA events on the library's stream, state resident in HBM.
`e2e` = the same metric for a whole `ba`-style job through the C ABI with host buffers:
gbp_cuda_init (H2D of the problem) + K x [gbp_cuda_iterate(1) with the device-side
metric copied back] + gbp_cuda_get_beliefs (D2H), timed on the host clock.

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_FACTOR = 896  # logical fp32/int32 tensor elements one sweep touches per factor
# What the packed device layout actually moves per edge slot and sweep (gbp_layout.h): potential 224 r,
# camera message 112 r + 112 w, landmark message 48 r + 48 w, state records 16 r + 16 w + 16 r  (+ 64 B of
# landmark belief gathered through L2).  The landmark messages live in a persisting L2 window, so the DRAM
# traffic measured in steady state is lower still (profiles/traffic_k_sweep.json: steady_state).
MOVED_BYTES_PER_SLOT = 224 + 112 + 112 + 48 + 48 + 16 + 16 + 16
ALGO_BYTES_PER_CAMERA = 504
ALGO_BYTES_PER_LANDMARK = 144
WORKLOAD = {"cameras": 1000, "landmarks": 100000, "obs_per_point": 10.5, "seed": 1234}
BA_PREROLL = 12  # sweeps of the ba.cpp schedule (prior weakening at iters 1,3,5,7,9) before anything is timed


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.ok = index, [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        nv = self.nv if self.ok else None
        while self.ok and not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((mhz, int(reasons), util))
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        mhz = sorted(s[0] for s in self.samples)
        seen = set()
        for _, r, _ in self.samples:
            for bit, nm in names.items():
                if r & bit:
                    seen.add(nm)
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(seen),
                "samples": len(self.samples)}


def build_problem(scale=1):
    from gbp_poplar_b200 import BALProblem, Setup
    bal = BALProblem.synthetic(WORKLOAD["cameras"] * scale, WORKLOAD["landmarks"] * scale, WORKLOAD["obs_per_point"],
                               WORKLOAD["seed"])
    return bal, Setup(bal)


def ba_preroll(engine):
    for it in range(BA_PREROLL):
        if (it + 1) % 2 == 0 and it < 10:
            engine.weaken_priors()
        engine.iterate(1)


def page_lock_problem(problem):
    """Page-locks (cudaHostRegister) the caller-side arrays of a gbp_problem that gbp_cuda_init copies to the device, so
    that the end-to-end job reads its inputs from pinned host memory as the benchmark contract asks (the library copies
    straight from the caller's pointers).  Returns the registered pointers (for page_unlock) and a description."""
    C_, L_, E_ = problem.n_keyframes, problem.n_points, problem.n_edges
    sizes = {"cam_ids": 4 * E_, "lmk_ids": 4 * E_, "measurements": 8 * E_, "meas_variances": 4 * E_,
             "cam_priors_eta": 24 * C_, "cam_priors_lambda": 144 * C_, "lmk_priors_eta": 12 * L_,
             "lmk_priors_lambda": 36 * L_, "cam_scaling": 4 * C_, "lmk_scaling": 4 * L_, "cam_weaken_flag": 4 * C_,
             "lmk_weaken_flag": 4 * L_, "active_flag": 4 * E_, "damping": 4 * E_, "damping_count": 4 * E_,
             "mu": 36 * E_, "oldmu": 36 * E_}
    done = []
    try:
        import ctypes
        import torch
        rt = torch.cuda.cudart()
        for field, nbytes in sizes.items():
            ptr = ctypes.cast(getattr(problem, field), ctypes.c_void_p).value
            if ptr and nbytes:
                err = rt.cudaHostRegister(ptr, nbytes, 0)
                if int(err) != 0:
                    raise RuntimeError(f"cudaHostRegister({field}) -> {err}")
                done.append(ptr)
        return done, "page-locked (cudaHostRegister on the caller's arrays)"
    except Exception as e:  # noqa: BLE001 -- the job then runs from pageable memory, and says so
        page_unlock(done)
        return [], f"pageable ({e!r})"


def page_unlock(ptrs):
    if ptrs:
        import torch
        rt = torch.cuda.cudart()
        for ptr in ptrs:
            rt.cudaHostUnregister(ptr)


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_run(setup, n_sweeps, budget_s, warmup=1):
    """Times the reference's codelet arithmetic on the host cores (oracle/_ref when it was built,
    else the bit-identical port) on the SAME problem; sweeps only, metric excluded."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    kind = "reference" if oracle_lib.available("reference") else "port"
    # every host core this process may run on, set explicitly: torch.distributed.run exports OMP_NUM_THREADS=1,
    # which would otherwise pin the reference to one thread
    eng = oracle_lib.OracleEngine(setup.problem, kind=kind, threads=host_cores())
    eng.iterate(warmup)
    done, ms = 0, 0.0
    t0 = time.time()
    while done < n_sweeps and (time.time() - t0) < budget_s:
        eng.iterate(1)
        ms += eng.last_ms()
        done += 1
    E = setup.problem.n_edges
    return {"value": E * done / (ms / 1e3), "unit": "factor-updates/s", "cores": eng.threads, "kind": kind,
            "sample": f"{done} full sweeps of the same {E}-factor problem after init, all host threads (OpenMP static "
                      f"over edges per BSP step), sweeps only", "ms_per_sweep": ms / max(done, 1)}, done, ms


PARITY_TENSORS = ["cam_beliefs_eta", "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda", "cam_messages_eta",
                  "lmk_messages_eta", "lmk_messages_lambda", "factor_potentials_eta", "damping", "damping_count"]


def state_digest(engine):
    """SHA-256 over this rank's beliefs, messages, potentials and damping state."""
    import hashlib
    h = hashlib.sha256()
    for name in PARITY_TENSORS:
        h.update(engine.get_tensor(name).tobytes())
    return h.hexdigest()


def parity_full_size(make_engine, eng_digest, n_sweeps_after_preroll, world, dist):
    """The timed handle against a SECOND handle of the same (sharded) problem that reaches the same sweep count through
    single-sweep calls with per-sweep metrics -- complete sweeps, plain call pattern, the path the parity tests pin
    to the oracle.  Every rank compares the SHA-256 of its own state; any difference on any rank fails."""
    ref = make_engine()
    ba_preroll(ref)
    for _ in range(n_sweeps_after_preroll):
        ref.iterate(1, stats=True)
    same = state_digest(ref) == eng_digest
    ref.close()
    if world > 1:
        flags = [None] * world
        dist.all_gather_object(flags, bool(same))
        return flags
    return [bool(same)]


def parity_small_graph(world, rank, local_rank, dist, with_oracle):
    """A graph the CPU oracle finishes in seconds (64 cameras / 6 k landmarks per rank), run by all ranks through the call
    pattern that is timed (single-sweep preroll, then blocks without metrics), every tensor of every rank compared bit
    for bit with the reference's codelets summing beliefs in the multi-GPU order (rank 0; the oracle is the checker)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gbp_poplar_b200 import BALProblem, GBPEngine, Setup, default_opts
    st = Setup(BALProblem.synthetic(64 * world, 6000 * world, 9.0, 11))
    opts = default_opts(device=local_rank)
    eng = GBPEngine.sharded(st.problem, opts) if world > 1 else GBPEngine(st.problem, opts)
    ba_preroll(eng)
    for n in (7, 1, 13, 30):
        eng.iterate(n)
    import shard_worker
    res = {t: eng.get_tensor(t) for t in shard_worker.DUMP}
    if world > 1:
        sh = eng.shard
        res.update(lmk_global=np.array(sh.lmk_global), edge_global=np.array(sh.edge_global),
                   cam_range=np.array([sh.cam_begin, sh.cam_end]), cam_bounds=np.array(sh.cam_bounds))
    else:
        p = st.problem
        res.update(lmk_global=np.arange(p.n_points), edge_global=np.arange(p.n_edges),
                   cam_range=np.array([0, p.n_keyframes]), cam_bounds=np.array([0, p.n_keyframes]))
    n_boundary = eng.shard.n_boundary_points if eng.shard else 0
    eng.close()
    ranks = [res]
    if world > 1:
        ranks = [None] * world
        dist.all_gather_object(ranks, res)
    if rank != 0:
        return None
    out = {"graph": {"cameras": st.problem.n_keyframes, "landmarks": st.problem.n_points, "factors": st.problem.n_edges,
                     "boundary_landmarks": n_boundary}, "sweeps": 12 + 7 + 1 + 13 + 30}
    if not with_oracle:
        out["status"] = "oracle not run (--no-cpu-baseline)"
        return out
    import common
    import oracle_lib
    from test_sharding_cpu import check_against_global
    kind = "reference" if oracle_lib.available("reference") else "port"
    ora = oracle_lib.OracleEngine(st.problem, kind=kind, threads=host_cores())
    if world > 1:
        ora.set_shard_bounds(np.array(ranks[0]["cam_bounds"]))
    else:
        ora.set_reduce_order(1)
    common.run_ba(ora, 12)
    ora.iterate(7 + 1 + 13 + 30)
    try:
        check_against_global(ranks, ora, st, exact=True)
        out["status"] = "ok"
    except AssertionError as e:
        out["status"] = f"MISMATCH: {e}"
    out["oracle"] = kind
    out["what"] = "every tensor of every rank bit-identical to the CPU oracle"
    return out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    bal, setup = build_problem(scale=max(args.gpus, 1))   # the graph our arm runs at this N
    E = setup.problem.n_edges
    warm = args.warmup
    base, done, ms = cpu_baseline_run(setup, args.steps, budget_s=600.0, warmup=warm)
    value = base["value"]
    line = {
        "impl": "reference", "metric": "factor_message_updates_per_sec", "value": value, "unit": "factor-updates/s",
        "n_gpus": args.gpus, "steps": done, "warmup": warm, "ms_per_step": ms / max(done, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "sweeps_per_sec": 1e3 * done / ms,
        "config": {"workload": "synthetic BAL 1k cameras / 100k landmarks / ~1M factors per GPU (configs[3]); at N>1 "
                               "one N-times larger graph (N=8 ~ configs[4])",
                   "cameras": bal.n_keyframes, "landmarks": bal.n_points, "factors": E,
                   "note": "reference codelets (gbp_codelets.cpp) on host cores; Poplar IPUModel is not installable"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "factor-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="config4", choices=["config4", "config5"],
                    help="config4 (default): BASELINE.json configs[3] per GPU, one N-times larger graph at N>1 (weak "
                         "scaling); config5: configs[4] exactly -- 10k cameras / 1M landmarks / ~10M factors in total, "
                         "partitioned over the N GPUs (strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the GBP hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from gbp_poplar_b200 import GBPEngine, default_opts

    bal, setup = build_problem(scale=world if args.workload == "config4" else 10)   # ONE graph
    E, Cn, Ln = setup.problem.n_edges, setup.problem.n_keyframes, setup.problem.n_points

    def make_engine():
        if world > 1:
            return GBPEngine.sharded(setup.problem, opts)
        return GBPEngine(setup.problem, opts)

    opts = default_opts(device=local_rank)
    t_init0 = time.time()
    eng = make_engine()
    init_s = time.time() - t_init0
    E_loc, C_loc, L_loc = eng.n_edges, eng.n_keyframes, eng.n_points   # this rank's shard (== global at N=1)
    n_boundary = eng.shard.n_boundary_points if eng.shard else 0
    exchange_mode = {"p2p": ("peer-to-peer stores" if os.environ.get("GBP_XCHG_PUSH", "0") not in ("", "0") else "peer-to-peer loads")
                            + " over NVLink (CUDA IPC) inside the belief-update kernel, step-tagged words, no collective",
                     "nccl": "NCCL all-gather", "none": "nothing"}[eng.exchange_mode()]
    ba_preroll(eng)

    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the timed region: K sweeps, device-timed on the library's stream
    eng.iterate(args.warmup)
    barrier()
    eng.iterate(args.steps)
    ms, launches = eng.last_timing()
    barrier()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = E * args.steps / (ms / 1e3)              # whole-job: all factors of the one graph

    # ---- per-kernel durations (CUDA events around every launch) for the roofline
    eng.set_profile(True)
    eng.iterate(args.steps)
    ms_prof, _ = eng.last_timing()
    ms_factor, ms_var = eng.last_kernel_times()
    sweep_ms_factor, _ = eng.last_sweep_times()      # per sweep: relinearising sweeps are a different workload
    eng.set_profile(False)
    stats = eng.eval()
    # ---- results of the timed handle, checked (at every N): see parity_full_size / parity_small_graph
    digest = state_digest(eng)
    parity_flags = parity_full_size(make_engine, digest, args.warmup + 2 * args.steps, world, dist if world > 1 else None)
    parity_small = parity_small_graph(world, rank, local_rank, dist if world > 1 else None, not args.no_cpu_baseline)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()
    if world > 1:
        # every rank samples its own GPU: one power-capped GPU slows the whole lock-step job
        try:
            per_rank = [None] * world
            dist.all_gather_object(per_rank, clocks)
            mhz = [c["sm_mhz"] for c in per_rank if c and c.get("sm_mhz")]
            clocks = dict(clocks)
            clocks["ranks"] = [{"sm_mhz": c.get("sm_mhz"), "reasons": c.get("reasons")} for c in per_rank if c]
            if mhz:
                clocks["sm_mhz"] = min(mhz)
            clocks["reasons"] = sorted(set(r for c in per_rank if c for r in (c.get("reasons") or [])))
        except Exception as e:  # never let the diagnostics break the measurement
            clocks = dict(clocks)
            clocks["ranks_error"] = repr(e)

    per_rank_kernels = None
    if world > 1:
        # factor / variable kernel time of every rank: the lock-step job runs at the slowest rank's pace and
        # the others show the difference as waiting time inside k_update_vars
        try:
            per_rank_kernels = [None] * world
            dist.all_gather_object(per_rank_kernels, [ms_factor / args.steps * 1e3, ms_var / args.steps * 1e3])
        except Exception as e:
            per_rank_kernels = repr(e)

    peak, peak_src = read_peaks()
    algo_bytes_factor_kernel = ALGO_BYTES_PER_FACTOR * E_loc   # one launch = this rank's factors
    t_factor = ms_factor / args.steps / 1e3
    achieved = algo_bytes_factor_kernel / t_factor / 1e9
    roofline = {
        "bound": "hbm", "kernel": ("k_sweep" if os.environ.get("GBP_SWEEP", "tma") == "cpasync" else "k_sweep_tma") + "<PREP=true,MSG=true,UPPER=false> (the last sweep of a call runs UPPER=true)", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
        "note": "achieved/frac use the contract figure (896 B per factor = every logical tensor element of the "
                "reference, SURVEY 8d); the packed layout moves fewer bytes, see achieved_moved / frac_moved",
        "moved_bytes_per_launch_model": MOVED_BYTES_PER_SLOT * (-(-E_loc // 32) * 32),
        "algorithmic_bytes_per_launch": algo_bytes_factor_kernel,
        "avg_launch_us": t_factor * 1e6, "kernel_share_of_step": ms_factor / max(ms_prof, 1e-9),
        "variable_kernel_avg_us": ms_var / args.steps * 1e3,
        "per_rank_kernel_us": per_rank_kernels,
        "whole_sweep_algorithmic_GBps": (algo_bytes_factor_kernel + ALGO_BYTES_PER_CAMERA * Cn +
                                         ALGO_BYTES_PER_LANDMARK * Ln) * args.steps / (ms / 1e3) / 1e9,
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic_k_sweep.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            tr = json.load(f)
        if tr.get("factors") == E_loc:
            roofline["traffic"] = tr.get("dram_bytes_per_launch")
            roofline["traffic_source"] = tr.get("source")
            ss = tr.get("steady_state") or {}
            if ss:  # the --set full capture flushes L2 before every replay; this one leaves the persisting window warm
                roofline["traffic_steady_state"] = ss.get("k_sweep_dram_bytes_per_launch")
                roofline["traffic_steady_state_source"] = ss.get("source")
    # Bytes actually moved, per class of launch.  A sweep in which the factors relinearise (all of them at once on a
    # fresh graph: one sweep in eleven) also rewrites the 224-byte potentials and runs ~2x the instructions, so it is
    # its own class: launches are classified by their measured duration (> 1.5 x the median), every class divides ITS
    # DRAM traffic (ncu, profiles/traffic_k_sweep.json) by ITS average duration, and frac_moved is the time-weighted
    # mean = total bytes moved / total kernel time / peak.
    import numpy as np
    t_us = np.asarray(sweep_ms_factor, dtype=np.float64) * 1e3
    med = float(np.median(t_us)) if t_us.size else 0.0
    relin = t_us > 1.5 * med
    model = roofline["moved_bytes_per_launch_model"]
    ss = (tr.get("steady_state") or {}) if os.path.exists(traffic_file) and tr.get("factors") == E_loc else {}
    b_plain = ss.get("k_sweep_dram_bytes_per_launch") or roofline["traffic"] or model
    b_relin = ss.get("k_sweep_relinearising_dram_bytes_per_launch") or (b_plain + 224 * E_loc)
    classes = {}
    for name, mask, nbytes in (("plain", ~relin, b_plain), ("relinearising", relin, b_relin)):
        if mask.any():
            t_cls = float(t_us[mask].mean())
            classes[name] = {"launches": int(mask.sum()), "avg_launch_us": t_cls, "dram_bytes_per_launch": nbytes,
                             "achieved": nbytes / (t_cls * 1e-6) / 1e9, "frac": nbytes / (t_cls * 1e-6) / 1e9 / peak}
    total_bytes = sum(c["launches"] * c["dram_bytes_per_launch"] for c in classes.values())
    total_s = float(t_us.sum()) * 1e-6
    roofline["classes"] = classes
    roofline["achieved_moved"] = total_bytes / total_s / 1e9 if total_s > 0 else None
    roofline["frac_moved"] = roofline["achieved_moved"] / peak if total_s > 0 else None
    roofline["frac_moved_what"] = ("sum over launch classes of (launches x measured DRAM bytes per launch) / total k_sweep time / peak; "
                                   "per-class figures in `classes`" + ("" if ss else " (no ncu traffic on record for this shard size: layout model)"))

    # ---- the opt-in contracted-FMA build of the sweep kernel (gbp_opts.fast_math), reported BESIDE the bit-faithful
    # default above, never instead of it: same workload, same call pattern, same timing
    fma = None
    try:
        opts_f = default_opts(device=local_rank, fast_math=1)
        eng_f = GBPEngine.sharded(setup.problem, opts_f) if world > 1 else GBPEngine(setup.problem, opts_f)
        ba_preroll(eng_f)
        eng_f.iterate(args.warmup)
        barrier()
        eng_f.iterate(args.steps)
        ms_f, _ = eng_f.last_timing()
        barrier()
        if world > 1:
            t = torch.tensor([ms_f], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_f = float(t.item())
        eng_f.set_profile(True)
        eng_f.iterate(args.steps)
        ms_factor_f, ms_var_f = eng_f.last_kernel_times()
        eng_f.set_profile(False)
        stats_f = eng_f.eval()
        eng_f.close()
        fma = {"value": E * args.steps / (ms_f / 1e3), "unit": "factor-updates/s", "ms_per_step": ms_f / args.steps,
               "k_sweep_avg_us": ms_factor_f / args.steps * 1e3, "final_reproj_px": stats_f["reproj_mean"],
               "final_reproj_px_bit_faithful": stats["reproj_mean"],
               "what": "gbp_opts.fast_math = 1: multiply-adds contracted (one rounding instead of two); NOT bit-comparable with "
                       "the reference, validated at the north-star tolerance (tests/test_fast_math_gpu.py)"}
    except Exception as e:  # the default measurement never depends on the optional build
        fma = {"error": repr(e)}

    # ---- e2e: a whole ba-style job through the C ABI with host buffers (rank-local problem)
    e2e = None
    eng.close()                                        # the e2e job below is a fresh one: nothing of the timed engine is kept
    if rank == 0 or world > 1:
        locked, host_buffers = page_lock_problem(setup.problem)   # untimed: the job's inputs sit in pinned host memory
        barrier()
        t0 = time.time()
        eng2 = make_engine()                           # H2D of the (rank's part of the) problem + LINEARISE_PROG
        t_init = time.time() - t0
        last = None
        for it in range(args.steps):
            if (it + 1) % 2 == 0 and it < 10:
                eng2.weaken_priors()
            last = eng2.iterate(1, stats=True)[0]      # D2H of the per-sweep metric
        t_loop = time.time() - t0 - t_init
        beliefs = eng2.get_beliefs()                   # READ_PROG: D2H beliefs + damping state
        wall = time.time() - t0
        # what gbp_cuda_init copies to the device: camera / landmark ids, measurements, variances (the optional per-edge
        # arrays are NULL = defaults here), the priors, scalings and weaken flags of both variable kinds
        h2d = 4 * (2 * E_loc + 2 * E_loc + E_loc + 42 * C_loc + 12 * L_loc + 2 * C_loc + 2 * L_loc)
        d2h = 24 * args.steps + sum(v.nbytes for v in beliefs.values())
        if world > 1:
            t = torch.tensor([float(h2d), float(d2h)], device="cuda")
            dist.all_reduce(t)
            h2d, d2h = float(t[0].item()), float(t[1].item())
        if world > 1:
            t = torch.tensor([wall], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        e2e = {"value": E * args.steps / wall, "unit": "factor-updates/s",
               "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "wall_s": wall, "init_s": t_init, "loop_s": t_loop, "read_s": wall - t_init - t_loop, "final_reproj_px": last["reproj_mean"] if last else None,
               "what": "gbp_cuda_init + K x gbp_cuda_iterate(1, stats) + gbp_cuda_get_beliefs, host clock",
               "host_buffers": host_buffers}
        eng2.close()
        page_unlock(locked)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_baseline_run(setup, 6, budget_s=40.0)

    if rank == 0:
        line = {
            "metric": "factor_message_updates_per_sec", "value": value, "unit": "factor-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak" if args.workload == "config4" else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "sweeps_per_sec": args.steps / (ms / 1e3) * 1.0,
            "config": {"workload": ("synthetic BAL 1k cameras / 100k landmarks / ~1M factors per GPU (configs[3]); at N>1 "
                                    "one N-times larger graph partitioned by camera range (N=8 ~ configs[4])")
                       if args.workload == "config4" else
                       "synthetic BAL 10k cameras / 1M landmarks / ~10M factors in total (configs[4]), partitioned by camera range",
                       "cameras": Cn, "landmarks": Ln, "factors": E, "per_gpu": False,
                       "parallelism": f"camera-range shards x{world}, boundary-landmark partials exchanged per sweep by "
                                      f"{exchange_mode} ({n_boundary} boundary landmarks in the whole graph, 96 B each per observing peer)"
                       if world > 1 else "single",
                       "rank0_shard": {"cameras": C_loc, "landmarks": L_loc, "factors": E_loc},
                       "cache": "per-sweep working set ~0.65 GB per GPU >> 126 MB L2 (no flush needed)",
                       "preroll": f"{BA_PREROLL} sweeps of the ba.cpp schedule incl. prior weakening, untimed",
                       "init_s": init_s},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "fma_mode": fma,
            "parity_n": {"status": "ok" if all(parity_flags) and (parity_small or {}).get("status") in ("ok", "oracle not run (--no-cpu-baseline)")
                         else "MISMATCH",
                         "full_size": {"ranks_identical": parity_flags,
                                       "what": "SHA-256 of every rank's beliefs / messages / potentials / damping state after the "
                                               "timed block calls == a second handle stepped one sweep per call with metrics"},
                         "small_graph_vs_oracle": parity_small},
            "clocks": clocks,
            "final": stats,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
