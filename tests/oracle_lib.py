"""Test-only loader for the CPU oracle libraries (oracle/).

  port      -> oracle/libgbp_oracle.so    (restated arithmetic, oracle/gbp_restated.hpp)
  reference -> oracle/_ref/libgbp_ref.so  (the reference's own codelet sources behind oracle/shim)

Both export the engine API of include/gbp_cuda.h under the prefix
`gbp_oracle_`, so tests drive them through gbp_poplar_b200.engine.GBPEngine.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from gbp_poplar_b200.engine import GBPEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_PATH = os.path.join(ROOT, "oracle", "libgbp_oracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libgbp_ref.so")
_libs = {}


def build(kind="port"):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port" if kind == "port" else "ref"])


def available(kind):
    return os.path.exists(PORT_PATH if kind == "port" else REF_PATH)


def load(kind="port"):
    if kind not in _libs:
        path = PORT_PATH if kind == "port" else REF_PATH
        if not os.path.exists(path):
            build(kind)
        lib = C.CDLL(path)
        lib.gbp_oracle_kind.restype = C.c_char_p
        lib.gbp_oracle_last_error.restype = C.c_char_p
        lib.gbp_oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        lib.gbp_oracle_set_reduce_order.argtypes = [C.c_void_p, C.c_int]
        lib.gbp_oracle_set_shard_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32]
        lib.gbp_oracle_commit_messages.argtypes = [C.c_void_p]
        lib.gbp_oracle_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        f32p = C.POINTER(C.c_float)
        lib.gbp_oracle_inv6x6.argtypes = [f32p, f32p]
        lib.gbp_oracle_inv3x3.argtypes = [f32p, f32p]
        lib.gbp_oracle_project.argtypes = [f32p] * 6
        assert lib.gbp_oracle_kind().decode() == ("port" if kind == "port" else "reference")
        _libs[kind] = lib
    return _libs[kind]


class OracleEngine(GBPEngine):
    """GBPEngine bound to an oracle library (same programs, CPU arithmetic)."""

    def __init__(self, problem, opts=None, kind="port", threads=None, keepalive=None):
        lib = load(kind)
        super().__init__(problem, opts, lib=lib, prefix="gbp_oracle_", keepalive=keepalive)
        self.kind = kind
        n = threads if threads is not None else lib.gbp_oracle_max_threads()
        lib.gbp_oracle_set_threads(self._h, int(n))
        self.threads = int(n)

    def set_reduce_order(self, mode):
        """0 = serial slot order (default), 1 = the CUDA path's tile order (bit-comparable)."""
        self._check(self._lib.gbp_oracle_set_reduce_order(self._h, int(mode)))

    def set_shard_bounds(self, bounds):
        """Belief summation order of the multi-GPU path for the camera-range partition `bounds` [world+1]."""
        b = np.ascontiguousarray(bounds, dtype=np.uint32)
        self._check(self._lib.gbp_oracle_set_shard_bounds(self._h, b.ctypes.data_as(C.POINTER(C.c_uint32)), b.size - 1))

    def commit_messages(self):
        self._check(self._lib.gbp_oracle_commit_messages(self._h))

    def last_ms(self):
        ms = C.c_double()
        self._lib.gbp_oracle_last_timing(self._h, C.byref(ms))
        return ms.value


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def inv6x6(A, kind="port"):
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(36)
    out = np.empty(36, np.float32)
    load(kind).gbp_oracle_inv6x6(_f(A), _f(out))
    return out.reshape(6, 6)


def inv3x3(A, kind="port"):
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(9)
    out = np.empty(9, np.float32)
    load(kind).gbp_oracle_inv3x3(_f(A), _f(out))
    return out.reshape(3, 3)


def project(x, p, K, kind="port"):
    x = np.ascontiguousarray(x, dtype=np.float32)
    p = np.ascontiguousarray(p, dtype=np.float32)
    K = np.ascontiguousarray(K, dtype=np.float32)
    hx, Jk, Jl = np.empty(2, np.float32), np.empty(12, np.float32), np.empty(6, np.float32)
    load(kind).gbp_oracle_project(_f(x), _f(p), _f(K), _f(hx), _f(Jk), _f(Jl))
    return hx, Jk.reshape(2, 6), Jl.reshape(2, 3)
