"""Host-side problem setup (thin ctypes wrappers over include/gbp_host.h).

Mirrors the host half of the reference's mains: `BALProblem`
(include/dataio.h:13-75) and the preparation ba.cpp:489-590 /
slam.cpp:489-597 performs before WRITE_PROG.  All work happens in the C++
library; this module only owns handles and exposes numpy views.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import GbpCliOptions, GbpProblem

MODE_BA = 0
MODE_SLAM = 1


def _check(rc, lib):
    if rc != 0:
        raise RuntimeError(f"gbp error {rc}: {lib.gbp_cuda_last_error().decode()}")


class _OwnedArray(np.ndarray):
    """ndarray view into library-owned memory that keeps its owner alive."""
    _owner = None


def _view(ptr, n, dtype, owner=None):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).view(_OwnedArray)
    a._owner = owner
    return a


class BALProblem:
    """Parsed BAL-format problem (BALProblem::LoadFile, ba/dataio.cpp:17-57)."""

    def __init__(self, handle):
        self._lib = _capi.load_library()
        self._h = handle
        c, l, e = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(self._lib.gbp_bal_dims(self._h, c, l, e), self._lib)
        self.n_keyframes, self.n_points, self.n_edges = c.value, l.value, e.value

    @classmethod
    def load(cls, path):
        lib = _capi.load_library()
        h = C.c_void_p()
        rc = lib.gbp_bal_load(str(path).encode(), C.byref(h))
        if rc != 0:
            raise FileNotFoundError(f"ERROR: unable to open file {path}")
        return cls(h)

    @classmethod
    def from_arrays(cls, intrinsics, cam_idx, lmk_idx, observations, cameras, points):
        lib = _capi.load_library()
        intr = np.ascontiguousarray(intrinsics, dtype=np.float64)
        ci = np.ascontiguousarray(cam_idx, dtype=np.uint32)
        li = np.ascontiguousarray(lmk_idx, dtype=np.uint32)
        ob = np.ascontiguousarray(observations, dtype=np.float64).reshape(-1)
        cams = np.ascontiguousarray(cameras, dtype=np.float64).reshape(-1)
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1)
        h = C.c_void_p()
        _check(lib.gbp_bal_from_arrays(cams.size // 6, pts.size // 3, ci.size,
                                       intr.ctypes.data_as(_capi.c_f64p), ci.ctypes.data_as(_capi.c_u32p),
                                       li.ctypes.data_as(_capi.c_u32p), ob.ctypes.data_as(_capi.c_f64p),
                                       cams.ctypes.data_as(_capi.c_f64p), pts.ctypes.data_as(_capi.c_f64p),
                                       C.byref(h)), lib)
        return cls(h)

    @classmethod
    def synthetic(cls, n_cameras, n_points, obs_per_point=10.0, seed=1234):
        """Synthetic BAL-format problem, deterministic in seed."""
        lib = _capi.load_library()
        h = C.c_void_p()
        _check(lib.gbp_synth_generate(n_cameras, n_points, float(obs_per_point), seed, C.byref(h)), lib)
        return cls(h)

    def save(self, path):
        _check(self._lib.gbp_bal_save(self._h, str(path).encode()), self._lib)

    def with_means(self, beliefs):
        """The optimised problem: a copy whose parameters are the means of `beliefs` (GBPEngine.get_beliefs())."""
        f = lambda k: np.ascontiguousarray(beliefs[k], dtype=np.float32).ctypes.data_as(_capi.c_f32p)
        h = C.c_void_p()
        _check(self._lib.gbp_bal_with_means(self._h, f("cam_beliefs_eta"), f("cam_beliefs_lambda"), f("lmk_beliefs_eta"),
                                            f("lmk_beliefs_lambda"), C.byref(h)), self._lib)
        return BALProblem(h)

    @property
    def camera_index(self):
        return _view(self._lib.gbp_bal_camera_index(self._h), self.n_edges, np.uint32, self)

    @property
    def point_index(self):
        return _view(self._lib.gbp_bal_point_index(self._h), self.n_edges, np.uint32, self)

    @property
    def observations(self):
        return _view(self._lib.gbp_bal_observations(self._h), 2 * self.n_edges, np.float64, self)

    @property
    def parameters(self):
        return _view(self._lib.gbp_bal_parameters(self._h), 6 * self.n_keyframes + 3 * self.n_points, np.float64, self)

    @property
    def intrinsics(self):
        return _view(self._lib.gbp_bal_intrinsics(self._h), 4, np.float64, self)

    def close(self):
        if self._h:
            self._lib.gbp_bal_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def cli_options(**kw):
    """gbp_cli_options with the reference defaults (ba/ba.cpp:400-465), overridden by kw."""
    lib = _capi.load_library()
    o = GbpCliOptions()
    lib.gbp_cli_options_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


class Setup:
    """Host arrays of one problem = what ba.cpp / slam.cpp stream in WRITE_PROG."""

    def __init__(self, bal, options=None, mode=MODE_BA):
        self._lib = _capi.load_library()
        self.bal = bal
        self.options = options if options is not None else cli_options()
        self.mode = mode
        self._h = C.c_void_p()
        _check(self._lib.gbp_setup_create(bal._h, C.byref(self.options), mode, C.byref(self._h)), self._lib)

    @property
    def problem(self):
        """The `gbp_problem` (ctypes struct, pointers owned by this Setup)."""
        return self._lib.gbp_setup_problem(self._h).contents

    def array(self, field):
        """numpy view of one gbp_problem array field."""
        p = self.problem
        C_, L_, E_ = p.n_keyframes, p.n_points, p.n_edges
        sizes = {
            "cam_ids": (E_, np.uint32), "lmk_ids": (E_, np.uint32), "measurements": (2 * E_, np.float32),
            "meas_variances": (E_, np.float32), "cam_priors_eta": (6 * C_, np.float32),
            "cam_priors_lambda": (36 * C_, np.float32), "lmk_priors_eta": (3 * L_, np.float32),
            "lmk_priors_lambda": (9 * L_, np.float32), "cam_scaling": (C_, np.float32),
            "lmk_scaling": (L_, np.float32), "cam_weaken_flag": (C_, np.uint32),
            "lmk_weaken_flag": (L_, np.uint32), "active_flag": (E_, np.uint32), "damping": (E_, np.float32),
            "damping_count": (E_, np.int32), "mu": (9 * E_, np.float32), "oldmu": (9 * E_, np.float32),
        }
        n, dt = sizes[field]
        ptr = getattr(p, field)
        if not ptr:  # an optional array the setup leaves NULL because it only holds the default (include/gbp_cuda.h)
            default = {"active_flag": 1, "damping_count": -15}.get(field, 0)
            return np.full(n, default, dtype=dt)
        return _view(ptr, n, dt, self)

    @property
    def K(self):
        return np.array(list(self.problem.K), dtype=np.float32)

    @property
    def data_counter(self):
        return self._lib.gbp_setup_data_counter(self._h)

    def next_keyframe(self, cam_beliefs_eta, cam_beliefs_lambda, cam_priors_eta, cam_priors_lambda,
                      lmk_priors_eta, lmk_priors_lambda):
        """update_flags + initialise_new_kf + damping_count reset (ba/slam.cpp:1020-1041).

        The four prior arrays (float32, as returned by READ_PRIORS) are updated
        in place.  Returns (n_new_landmarks, damping_count array)."""
        E_ = self.problem.n_edges
        dc = np.empty(E_, dtype=np.int32)
        n_new = C.c_int()
        f = lambda a: a.ctypes.data_as(_capi.c_f32p)
        for a in (cam_beliefs_eta, cam_beliefs_lambda, cam_priors_eta, cam_priors_lambda, lmk_priors_eta,
                  lmk_priors_lambda):
            assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
        _check(self._lib.gbp_setup_next_keyframe(self._h, f(cam_beliefs_eta), f(cam_beliefs_lambda),
                                                 f(cam_priors_eta), f(cam_priors_lambda), f(lmk_priors_eta),
                                                 f(lmk_priors_lambda), dc.ctypes.data_as(_capi.c_i32p),
                                                 C.byref(n_new)), self._lib)
        return n_new.value, dc

    def close(self):
        if self._h:
            self._lib.gbp_setup_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Shard:
    """Rank-local sub-problem of a camera-range partition (gbp_shard_build, include/gbp_cuda.h).

    Replaces the reference's placement of variables and factors on the tiles of several
    IPUs (ba/ba.cpp:617-631,717-753,795-834).  Pure host code; `owner` keeps the global
    problem's arrays alive."""

    def __init__(self, problem, world, rank, owner=None, handle=None, view=False):
        self._lib = _capi.load_library()
        self._owner = owner
        self._owned = handle is None
        if handle is None:
            h = C.c_void_p()
            build = self._lib.gbp_shard_build_view if view else self._lib.gbp_shard_build
            _check(build(C.byref(problem), world, rank, C.byref(h)), self._lib)
            handle = h
        self._h = handle
        pl = self._lib.gbp_shard_get_plan(self._h).contents
        self.world, self.rank = pl.world, pl.rank
        self.cam_begin, self.cam_end = pl.cam_begin, pl.cam_end
        self.n_local_edges, self.n_local_points = pl.n_local_edges, pl.n_local_points
        self.n_boundary_points = pl.n_boundary_points
        self.n_boundary_local = self._lib.gbp_shard_n_boundary_local(self._h)
        self.n_active_global = self._lib.gbp_shard_n_active_global(self._h)

    @property
    def problem(self):
        return self._lib.gbp_shard_problem(self._h).contents

    @property
    def lmk_global(self):
        return _view(self._lib.gbp_shard_lmk_global(self._h), self.n_local_points, np.uint32, self)

    @property
    def edge_global(self):
        return _view(self._lib.gbp_shard_edge_global(self._h), self.n_local_edges, np.uint32, self)

    @property
    def boundary_local(self):
        return _view(self._lib.gbp_shard_boundary_local(self._h), self.n_boundary_local, np.uint32, self)

    @property
    def boundary_slot(self):
        return _view(self._lib.gbp_shard_boundary_slot(self._h), self.n_boundary_local, np.uint32, self)

    @property
    def boundary_ranks(self):
        """[n_boundary_local] bit r = rank r observes the boundary landmark (contributes a partial sum)."""
        return _view(self._lib.gbp_shard_boundary_ranks(self._h), self.n_boundary_local, np.uint32, self)

    @property
    def cam_bounds(self):
        return _view(self._lib.gbp_shard_cam_bounds(self._h), self.world + 1, np.uint32, self)

    def close(self):
        if self._h and self._owned:
            self._lib.gbp_shard_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def problem_from_arrays(arrays, K):
    """Build a gbp_problem struct from a dict of numpy arrays (kept alive by the caller)."""
    p = GbpProblem()
    p.n_keyframes = arrays["cam_scaling"].size
    p.n_points = arrays["lmk_scaling"].size
    p.n_edges = arrays["cam_ids"].size
    for i in range(9):
        p.K[i] = float(K[i])
    tmap = {np.dtype(np.float32): _capi.c_f32p, np.dtype(np.uint32): _capi.c_u32p, np.dtype(np.int32): _capi.c_i32p}
    for name, _ in GbpProblem._fields_:
        if name in ("n_keyframes", "n_points", "n_edges", "K"):
            continue
        a = arrays.get(name)
        if a is None:
            continue
        assert a.flags["C_CONTIGUOUS"], name
        setattr(p, name, a.ctypes.data_as(tmap[a.dtype]))
    return p
