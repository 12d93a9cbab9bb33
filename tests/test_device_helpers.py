"""The device math helpers (gbp_math.cuh: inv3x3, inv6x6, so3exp + the projection Jacobians) run on the GPU through the
library's self-test entry points and pinned, bit for bit, to the golden vectors generated from the reference's own
ba/matlib.cpp / ba/bafuncs.cpp (tests/golden/golden_helpers.npz, made by tests/golden/make_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from gbp_poplar_b200 import _capi

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _f(a):
    return a.ctypes.data_as(_capi.c_f32p)


@pytest.fixture(scope="module")
def z():
    return np.load(os.path.join(G, "golden_helpers.npz"))


def test_inv6x6_bit_exact(z):
    lib = _capi.load_library()
    A = np.ascontiguousarray(z["A6"], np.float32)
    out = np.empty_like(A)
    assert lib.gbp_cuda_test_inv6x6(_f(A), _f(out), len(A)) == 0, lib.gbp_cuda_last_error()
    assert out.tobytes() == np.ascontiguousarray(z["inv6"], np.float32).tobytes()
    # and it is an inverse (fp32 LDL^T of a moderately conditioned matrix)
    err = np.abs(out.astype(np.float64) @ A.astype(np.float64) - np.eye(6)).max()
    assert err < 5e-3, err


def test_inv3x3_bit_exact(z):
    lib = _capi.load_library()
    A = np.ascontiguousarray(z["A3"], np.float32)
    out = np.empty_like(A)
    assert lib.gbp_cuda_test_inv3x3(_f(A), _f(out), len(A)) == 0, lib.gbp_cuda_last_error()
    assert out.tobytes() == np.ascontiguousarray(z["inv3"], np.float32).tobytes()


def test_projection_and_jacobians_bit_exact(z):
    """so3exp -> hfunc -> Jac: h(x), d h / d pose (2x6) and d h / d landmark (2x3)."""
    lib = _capi.load_library()
    X = np.ascontiguousarray(z["X"], np.float32)
    P = np.ascontiguousarray(z["P"], np.float32)
    K = np.ascontiguousarray(z["K"], np.float32)
    n = len(X)
    hx, jk, jl = np.empty((n, 2), np.float32), np.empty((n, 12), np.float32), np.empty((n, 6), np.float32)
    assert lib.gbp_cuda_test_project(_f(X), _f(P), _f(K), _f(hx), _f(jk), _f(jl), n) == 0, lib.gbp_cuda_last_error()
    assert hx.tobytes() == np.ascontiguousarray(z["hx"], np.float32).tobytes()
    assert jk.tobytes() == np.ascontiguousarray(z["Jkf"], np.float32).tobytes()
    assert jl.tobytes() == np.ascontiguousarray(z["Jlmk"], np.float32).tobytes()


def test_empty_batches_are_fine():
    lib = _capi.load_library()
    a = np.zeros(36, np.float32)
    assert lib.gbp_cuda_test_inv6x6(_f(a), _f(a), 0) == 0
