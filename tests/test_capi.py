"""The C-ABI library: loads without a GPU, exports every symbol include/*.h
declares, fails loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import common
from gbp_poplar_b200 import _capi
from gbp_poplar_b200._capi import GbpShardPlan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _declared_in_headers():
    names = set()
    for h in ("gbp_cuda.h", "gbp_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(gbp_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = _capi.load_library()
    declared = _declared_in_headers()
    assert len(declared) > 40
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    # the ctypes mirror covers the same set
    assert set(_capi.declared_symbols()) == declared


def test_version_and_defaults():
    lib = _capi.load_library()
    assert b"sm_100a" in lib.gbp_cuda_version()
    o = _capi.GbpOpts()
    lib.gbp_opts_default(C.byref(o))
    assert (o.maxeta_damping, o.num_undamped_iters, o.min_linear_iters) == (pytest.approx(0.4), 8, 10)
    assert o.dmu_threshold == pytest.approx(3e-3) and o.Nstds == pytest.approx(2.5)


@pytest.mark.skipif(_has_gpu(), reason="a GPU is present")
def test_init_fails_loudly_without_a_gpu():
    from gbp_poplar_b200 import GBPEngine
    st = common.make_setup("fr2robot2")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GBPEngine(st.problem)


def test_null_arguments_are_rejected():
    lib = _capi.load_library()
    fns = _capi.bind_engine(lib, "gbp_cuda_")
    h = C.c_void_p()
    assert fns["init"](None, None, C.byref(h)) == -1
    assert fns["iterate"](None, 1, None) == -1
    assert fns["get_tensor"](None, b"mu", None, 0) == -1
    assert fns["add_keyframe"](None, None, None, None, None, None, None, None, None) == -1
    assert lib.gbp_cuda_add_keyframe_device(None, 2, 5, None) == -1


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_shard_partitions_cameras(world):
    lib = _capi.load_library()
    st = common.make_setup("fr1xyz")
    p = st.problem
    cam = st.array("cam_ids")
    prev_end, tot_edges = 0, 0
    for r in range(world):
        plan = GbpShardPlan()
        assert lib.gbp_cuda_plan_shard(C.byref(p), world, r, C.byref(plan)) == 0
        assert plan.cam_begin == prev_end and plan.cam_end >= plan.cam_begin
        prev_end = plan.cam_end
        n = int(((cam >= plan.cam_begin) & (cam < plan.cam_end)).sum())
        assert n == plan.n_local_edges
        tot_edges += n
        if world > 1:
            assert n < 2.0 * p.n_edges / world  # balanced by edge count
        if world == 1:
            assert plan.n_boundary_points == 0
    assert prev_end == p.n_keyframes and tot_edges == p.n_edges


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected before any CUDA call (same codes with or without a device)."""
    lib = _capi.load_library()
    h = C.c_void_p()
    assert lib.gbp_cuda_init(None, None, C.byref(h)) == -1                      # GBP_ERR_ARG
    p = _capi.GbpProblem()
    p.n_edges = 5                                                               # arrays missing
    assert lib.gbp_cuda_init(C.byref(p), None, C.byref(h)) == -1
    assert b"null required array" in lib.gbp_cuda_last_error()
    assert lib.gbp_cuda_free(None) == 0
    assert lib.gbp_cuda_iterate(None, 1, None) == -1
    n, why = C.c_int(), C.c_int()
    assert lib.gbp_cuda_iterate_until(None, 10, 5, 1e-3, 2.0, None, C.byref(n), C.byref(why)) == -1
    st = common.make_setup("fr2robot2")
    plan = GbpShardPlan()
    assert lib.gbp_cuda_plan_shard(C.byref(st.problem), 0, 0, C.byref(plan)) == -1
    assert lib.gbp_cuda_plan_shard(C.byref(st.problem), 4, 4, C.byref(plan)) == -1
    assert lib.gbp_cuda_exchange_mode(None) == 0
