"""GBPEngine: Python mirror of the reference's engine/program interface.

The reference host loop drives its device with `engine.run(PROG)` over seven
programs (ba/ba.cpp:925-934, ba/slam.cpp:937-948).  Each method below is the
call a maintainer would make instead, forwarding to the C ABI of
include/gbp_cuda.h:

    WRITE_PROG + LINEARISE_PROG -> GBPEngine(problem)        (gbp_cuda_init)
    GBP_PROG                    -> engine.iterate(n)         (gbp_cuda_iterate)
    WEAKEN_PRIORS               -> engine.weaken_priors()
    READ_PROG                   -> engine.get_beliefs()
    READ_PRIORS                 -> engine.get_priors()
    NEW_KEYFRAME                -> engine.add_keyframe(...)

The class is generic over (library, symbol prefix) only so that the test-suite
can drive the CPU oracle through the very same code; the product always binds
libgbp_cuda.so / "gbp_cuda_" and raises if that library is missing.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import GbpIterStats, GbpOpts

_F32 = np.float32
_TENSOR_DTYPES = {
    "damping_count": np.int32, "active_flag": np.uint32, "robust_flag": np.uint32,
    "cam_weaken_flag": np.uint32, "lmk_weaken_flag": np.uint32,
}
TENSOR_NAMES = [
    "cam_beliefs_eta", "cam_beliefs_lambda", "lmk_beliefs_eta", "lmk_beliefs_lambda",
    "cam_messages_eta", "cam_messages_lambda", "lmk_messages_eta", "lmk_messages_lambda",
    "pcam_messages_eta", "pcam_messages_lambda", "plmk_messages_eta", "plmk_messages_lambda",
    "factor_potentials_eta", "factor_potentials_lambda", "damping", "damping_count", "mu", "oldmu", "dmu",
    "active_flag", "robust_flag", "measurements", "meas_variances", "cam_scaling", "lmk_scaling",
    "cam_weaken_flag", "lmk_weaken_flag",
]


def default_opts(**kw):
    """gbp_opts with the hyper-parameters of ba/gbp_codelets.cpp:11-16 (gbp_opts_default: the library's own
    defaults, including its environment overrides)."""
    o = GbpOpts()
    _capi.load_library().gbp_opts_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


def stats_to_dict(s):
    return {"reproj_mean": s.reproj_mean, "cost": s.cost, "n_relins": s.n_relins, "n_robust": s.n_robust,
            "n_active": s.n_active}


class GBPEngine:
    _uid_cache = {}

    def __init__(self, problem, opts=None, lib=None, prefix="gbp_cuda_", keepalive=None, shard=None, adopt=None):
        """shard = (world, rank, nccl_unique_id bytes): this rank's part of the GLOBAL `problem`
        (gbp_cuda_init_shard); every program then includes the boundary-landmark exchange and is
        collective over the ranks.  See GBPEngine.sharded().  adopt = (handle, world, rank): wrap a handle
        that gbp_cuda_init_group built (GBPGroup owns it)."""
        if lib is None:
            lib = _capi.load_library()
        self._lib = lib
        self._prefix = prefix
        self._f = _capi.bind_engine(lib, prefix)
        self._keepalive = keepalive
        self.opts = opts if opts is not None else default_opts()
        self._h = C.c_void_p()
        self.shard = None
        self._borrowed = adopt is not None
        if adopt is not None:
            handle, world, rank = adopt
            self._h = C.c_void_p(handle)
            if world > 1:
                from .host import Shard
                self.shard = Shard(None, world, rank, owner=self, handle=C.c_void_p(lib.gbp_cuda_shard_info(self._h)))
        elif shard is None:
            self._check(self._f["init"](C.byref(problem), C.byref(self.opts), C.byref(self._h)))
        else:
            world, rank, uid = shard
            buf = C.create_string_buffer(bytes(uid), 128)
            self._check(lib.gbp_cuda_init_shard(C.byref(problem), C.byref(self.opts), world, rank, buf,
                                                C.byref(self._h)))
            if world > 1:
                from .host import Shard
                self.shard = Shard(None, world, rank, owner=self, handle=C.c_void_p(lib.gbp_cuda_shard_info(self._h)))
        c, l, e, mk, ml = (C.c_uint32() for _ in range(5))
        self._check(self._f["dims"](self._h, c, l, e, mk, ml))
        self.n_keyframes, self.n_points, self.n_edges = c.value, l.value, e.value
        self.max_nkfedges, self.max_nlmkedges = mk.value, ml.value

    @classmethod
    def sharded(cls, problem, opts=None, group=None):
        """One rank of a multi-GPU run under torch.distributed (one process per GPU): rank 0 creates
        the NCCL unique id, torch.distributed broadcasts it, every rank builds its shard."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            return cls(problem, opts)
        lib = _capi.load_library()
        key = (id(group), world, rank)
        if key in cls._uid_cache:  # same ranks again: the library reuses the communicator of this id
            return cls(problem, opts, shard=(world, rank, cls._uid_cache[key]))
        uid = C.create_string_buffer(128)
        if rank == 0:
            rc = lib.gbp_cuda_nccl_unique_id(uid)
            if rc != 0:
                raise RuntimeError(f"gbp_cuda_nccl_unique_id failed: {lib.gbp_cuda_last_error().decode()}")
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.frombuffer(bytearray(uid.raw), dtype=torch.uint8).to(dev)
        dist.broadcast(t, src=0, group=group)
        cls._uid_cache[key] = bytes(t.cpu().numpy().tobytes())
        return cls(problem, opts, shard=(world, rank, cls._uid_cache[key]))

    # -- plumbing ---------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = ""
            try:
                fn = getattr(self._lib, self._prefix + "last_error")
                fn.restype = C.c_char_p
                msg = fn().decode()
            except AttributeError:
                pass
            raise RuntimeError(f"{self._prefix}* failed with code {rc}: {msg}")

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h and not self._borrowed:
            self._f["free"](self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- programs ---------------------------------------------------------
    def weaken_priors(self):
        self._check(self._f["weaken_priors"](self._h))

    def iterate(self, n_sweeps=1, stats=False):
        if stats:
            arr = (GbpIterStats * n_sweeps)()
            self._check(self._f["iterate"](self._h, n_sweeps, arr))
            return [stats_to_dict(s) for s in arr]
        self._check(self._f["iterate"](self._h, n_sweeps, None))
        return None

    def iterate_until(self, max_sweeps, check_every=10, rel_tol=1e-4, diverge_factor=2.0):
        """Sweeps until the mean reprojection error stalls (relative improvement < rel_tol over `check_every`
        sweeps) or exceeds diverge_factor x its running minimum.  Returns (per-sweep stats, reason) with reason in
        {"max_sweeps", "converged", "diverged"}."""
        arr = (GbpIterStats * max(max_sweeps, 1))()
        n, why = C.c_int(), C.c_int()
        self._check(self._lib.gbp_cuda_iterate_until(self._h, max_sweeps, check_every, rel_tol, diverge_factor, arr,
                                                     C.byref(n), C.byref(why)))
        return [stats_to_dict(arr[i]) for i in range(n.value)], ("max_sweeps", "converged", "diverged")[why.value]

    def eval(self):
        s = GbpIterStats()
        self._check(self._f["eval"](self._h, C.byref(s)))
        return stats_to_dict(s)

    def get_beliefs(self):
        Cn, Ln, En = self.n_keyframes, self.n_points, self.n_edges
        out = {
            "cam_beliefs_eta": np.empty(6 * Cn, _F32), "cam_beliefs_lambda": np.empty(36 * Cn, _F32),
            "lmk_beliefs_eta": np.empty(3 * Ln, _F32), "lmk_beliefs_lambda": np.empty(9 * Ln, _F32),
            "damping": np.empty(En, _F32), "damping_count": np.empty(En, np.int32),
            "robust_flag": np.empty(En, np.uint32),
        }
        f = lambda k: out[k].ctypes.data_as(_capi.c_f32p)
        self._check(self._f["get_beliefs"](self._h, f("cam_beliefs_eta"), f("cam_beliefs_lambda"),
                                           f("lmk_beliefs_eta"), f("lmk_beliefs_lambda"), f("damping"),
                                           out["damping_count"].ctypes.data_as(_capi.c_i32p),
                                           out["robust_flag"].ctypes.data_as(_capi.c_u32p)))
        return out

    def get_priors(self):
        Cn, Ln = self.n_keyframes, self.n_points
        out = {"cam_priors_eta": np.empty(6 * Cn, _F32), "cam_priors_lambda": np.empty(36 * Cn, _F32),
               "lmk_priors_eta": np.empty(3 * Ln, _F32), "lmk_priors_lambda": np.empty(9 * Ln, _F32)}
        f = lambda k: out[k].ctypes.data_as(_capi.c_f32p)
        self._check(self._f["get_priors"](self._h, f("cam_priors_eta"), f("cam_priors_lambda"),
                                          f("lmk_priors_eta"), f("lmk_priors_lambda")))
        return out

    def add_keyframe(self, damping_count, cam_prior_eta, cam_prior_lambda, lmk_prior_eta, lmk_prior_lambda,
                     active_flag, cam_weaken_flag, lmk_weaken_flag):
        def p(a, dt, ct):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(ct)
        keep = []
        self._check(self._f["add_keyframe"](
            self._h, p(damping_count, np.int32, _capi.c_i32p), p(cam_prior_eta, _F32, _capi.c_f32p),
            p(cam_prior_lambda, _F32, _capi.c_f32p), p(lmk_prior_eta, _F32, _capi.c_f32p),
            p(lmk_prior_lambda, _F32, _capi.c_f32p), p(active_flag, np.uint32, _capi.c_u32p),
            p(cam_weaken_flag, np.uint32, _capi.c_u32p), p(lmk_weaken_flag, np.uint32, _capi.c_u32p)))

    def add_keyframe_device(self, new_cam, steps=5):
        """The whole insertion of slam.cpp:1020-1046 (update_flags + initialise_new_kf + NEW_KEYFRAME)
        on the device; returns the number of newly observed landmarks.  CUDA library only."""
        n_new = C.c_int()
        self._check(self._lib.gbp_cuda_add_keyframe_device(self._h, int(new_cam), int(steps), C.byref(n_new)))
        return n_new.value

    # -- codelet-level entry points (Execute(cs_*)) -------------------------
    def relinearise_factors(self):
        self._check(self._f["relinearise_factors"](self._h))

    def prep_messages(self):
        self._check(self._f["prep_messages"](self._h))

    def compute_messages(self):
        self._check(self._f["compute_messages"](self._h))

    def update_beliefs(self):
        self._check(self._f["update_beliefs"](self._h))

    def weaken_prior_vertices(self):
        self._check(self._f["weaken_prior_vertices"](self._h))

    # -- tensors by reference name ------------------------------------------
    def tensor_nbytes(self, name):
        n = C.c_size_t()
        self._check(self._f["tensor_nbytes"](self._h, name.encode(), C.byref(n)))
        return n.value

    def get_tensor(self, name):
        n = self.tensor_nbytes(name)
        dt = _TENSOR_DTYPES.get(name, _F32)
        out = np.empty(n // 4, dtype=dt)
        self._check(self._f["get_tensor"](self._h, name.encode(), out.ctypes.data_as(C.c_void_p), n))
        return out

    def set_tensor(self, name, arr):
        dt = _TENSOR_DTYPES.get(name, _F32)
        a = np.ascontiguousarray(arr, dtype=dt).reshape(-1)
        self._check(self._f["set_tensor"](self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.nbytes))

    def snapshot(self, names=None):
        return {n: self.get_tensor(n) for n in (names or TENSOR_NAMES)}

    def restore(self, snap):
        for n, a in snap.items():
            self.set_tensor(n, a)

    # -- CUDA-only conveniences -----------------------------------------------
    def last_timing(self):
        ms = C.c_float()
        k = C.c_uint64()
        self._check(self._lib.gbp_cuda_last_timing(self._h, C.byref(ms), C.byref(k)))
        return ms.value, k.value

    def set_profile(self, enabled):
        self._check(self._lib.gbp_cuda_set_profile(self._h, int(bool(enabled))))

    def last_kernel_times(self):
        """(ms in the factor kernel, ms in the variable kernel) of the last iterate() under set_profile(True)."""
        a, b = C.c_float(), C.c_float()
        self._check(self._lib.gbp_cuda_last_kernel_times(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_sweep_times(self):
        """Per sweep of the last iterate() under set_profile(True): (factor-kernel ms array, variable-kernel ms array)."""
        n = C.c_int()
        self._check(self._lib.gbp_cuda_last_sweep_times(self._h, None, None, 0, C.byref(n)))
        a, b = np.empty(n.value, _F32), np.empty(n.value, _F32)
        self._check(self._lib.gbp_cuda_last_sweep_times(self._h, a.ctypes.data_as(_capi.c_f32p), b.ctypes.data_as(_capi.c_f32p),
                                                        n.value, C.byref(n)))
        return a, b

    def exchange_mode(self):
        """'none' | 'nccl' | 'p2p': how a sharded engine exchanges boundary-landmark partials."""
        return ("none", "nccl", "p2p")[self._lib.gbp_cuda_exchange_mode(self._h)]

    def iterate_async(self, n_sweeps):
        self._check(self._lib.gbp_cuda_iterate_async(self._h, n_sweeps))

    def synchronize(self):
        self._check(self._lib.gbp_cuda_synchronize(self._h))


class GBPGroup:
    """All ranks of a multi-GPU run in ONE process (gbp_cuda_init_group) -- the shape of the reference's own
    multi-chip mode, one host program driving 2^k IPUs (--ipus, ba/ba.cpp:617-631).  `devices[r]` is the CUDA
    device of rank r; ordinals may repeat (shards sharing one GPU: the multi-rank protocol on a one-GPU box).
    Programs that exchange boundary partials go through the group; `ranks[r]` is a GBPEngine view of rank r's
    handle for the per-rank read-backs (get_tensor, get_beliefs, shard maps)."""

    def __init__(self, problem, world, devices=None, opts=None):
        self._lib = _capi.load_library()
        self.opts = opts if opts is not None else default_opts()
        self.world = int(world)
        dev = None
        if devices is not None:
            assert len(devices) == self.world
            dev = (C.c_int * self.world)(*[int(d) for d in devices])
        self._hs = (C.c_void_p * self.world)()
        rc = self._lib.gbp_cuda_init_group(C.byref(problem), C.byref(self.opts), self.world, dev, self._hs)
        if rc != 0:
            raise RuntimeError(f"gbp_cuda_init_group failed with code {rc}: {self._lib.gbp_cuda_last_error().decode()}")
        self.ranks = [GBPEngine(None, self.opts, adopt=(self._hs[r], self.world, r)) for r in range(self.world)]

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"gbp_cuda_group_* failed with code {rc}: {self._lib.gbp_cuda_last_error().decode()}")

    def iterate(self, n_sweeps=1, stats=False):
        if stats:
            arr = (GbpIterStats * n_sweeps)()
            self._check(self._lib.gbp_cuda_group_iterate(self._hs, self.world, n_sweeps, arr))
            return [stats_to_dict(s) for s in arr]
        self._check(self._lib.gbp_cuda_group_iterate(self._hs, self.world, n_sweeps, None))
        return None

    def weaken_priors(self):
        self._check(self._lib.gbp_cuda_group_weaken_priors(self._hs, self.world))

    def eval(self):
        s = GbpIterStats()
        self._check(self._lib.gbp_cuda_group_eval(self._hs, self.world, C.byref(s)))
        return stats_to_dict(s)

    def last_timing(self):
        """(max over ranks of the device time of the last iterate, kernels launched by all ranks)."""
        t = [r.last_timing() for r in self.ranks]
        return max(x[0] for x in t), sum(x[1] for x in t)

    def close(self):
        if self._hs is not None:
            for r in self.ranks:
                r.close()
            self._lib.gbp_cuda_group_free(self._hs, self.world)
            self._hs = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
