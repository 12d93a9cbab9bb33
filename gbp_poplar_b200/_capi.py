"""ctypes mirror of include/gbp_cuda.h and include/gbp_host.h.

Only declarations live here: structure layouts and function prototypes of the
C-ABI shared library (gbp_poplar_b200/libgbp_cuda.so).  Nothing in this module
computes anything; PyTorch is not involved.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GBP_CUDA_LIB selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("GBP_CUDA_LIB") or os.path.join(_HERE, "libgbp_cuda.so")

c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)
c_i32p = C.POINTER(C.c_int32)


class GbpProblem(C.Structure):
    """struct gbp_problem (include/gbp_cuda.h)."""

    _fields_ = [
        ("n_keyframes", C.c_uint32),
        ("n_points", C.c_uint32),
        ("n_edges", C.c_uint32),
        ("cam_ids", c_u32p),
        ("lmk_ids", c_u32p),
        ("measurements", c_f32p),
        ("meas_variances", c_f32p),
        ("K", C.c_float * 9),
        ("cam_priors_eta", c_f32p),
        ("cam_priors_lambda", c_f32p),
        ("lmk_priors_eta", c_f32p),
        ("lmk_priors_lambda", c_f32p),
        ("cam_scaling", c_f32p),
        ("lmk_scaling", c_f32p),
        ("cam_weaken_flag", c_u32p),
        ("lmk_weaken_flag", c_u32p),
        ("active_flag", c_u32p),
        ("damping", c_f32p),
        ("damping_count", c_i32p),
        ("mu", c_f32p),
        ("oldmu", c_f32p),
    ]


class GbpOpts(C.Structure):
    """struct gbp_opts (include/gbp_cuda.h)."""

    _fields_ = [
        ("device", C.c_int),
        ("maxeta_damping", C.c_float),
        ("num_undamped_iters", C.c_int),
        ("dmu_threshold", C.c_float),
        ("min_linear_iters", C.c_int),
        ("Nstds", C.c_float),
        ("use_cuda_graph", C.c_int),
        ("store_full_messages", C.c_int),
        ("exchange", C.c_int),
        ("relin_mode", C.c_int),
        ("fast_math", C.c_int),
        ("reserved", C.c_int * 3),
    ]


class GbpIterStats(C.Structure):
    """struct gbp_iter_stats (include/gbp_cuda.h)."""

    _fields_ = [
        ("reproj_mean", C.c_float),
        ("cost", C.c_float),
        ("n_relins", C.c_uint32),
        ("n_robust", C.c_uint32),
        ("n_active", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class GbpShardPlan(C.Structure):
    """struct gbp_shard_plan (include/gbp_cuda.h)."""

    _fields_ = [
        ("world", C.c_uint32),
        ("rank", C.c_uint32),
        ("cam_begin", C.c_uint32),
        ("cam_end", C.c_uint32),
        ("n_local_edges", C.c_uint32),
        ("n_local_points", C.c_uint32),
        ("n_boundary_points", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class GbpCliOptions(C.Structure):
    """struct gbp_cli_options (include/gbp_host.h)."""

    _fields_ = [
        ("n_iters", C.c_int),
        ("iters_between_kfs", C.c_int),
        ("n_ipus", C.c_int),
        ("cams_per_tile", C.c_int),
        ("profile", C.c_int),
        ("transnoise", C.c_float),
        ("rotnoise", C.c_float),
        ("lmktrans_noise", C.c_float),
        ("av_depth_on", C.c_int),
        ("av_depth", C.c_float),
        ("reproj_meas_var", C.c_float),
        ("prior_std_weaker_factor", C.c_float),
        ("first_cam_prior_std", C.c_float),
        ("steps", C.c_float),
        ("iters_before_damping", C.c_int),
        ("verbose", C.c_int),
        ("noise_seed", C.c_uint32),
    ]


# Engine-level entry points shared (same signatures) by libgbp_cuda.so
# (prefix "gbp_cuda_") and by the test-only oracle libraries (prefix
# "gbp_oracle_").  name -> (restype, argtypes) with the handle as void*.
ENGINE_API = {
    "init": (C.c_int, [C.POINTER(GbpProblem), C.POINTER(GbpOpts), C.POINTER(C.c_void_p)]),
    "free": (C.c_int, [C.c_void_p]),
    "weaken_priors": (C.c_int, [C.c_void_p]),
    "iterate": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(GbpIterStats)]),
    "eval": (C.c_int, [C.c_void_p, C.POINTER(GbpIterStats)]),
    "get_beliefs": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_u32p]),
    "get_priors": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p]),
    "add_keyframe": (C.c_int, [C.c_void_p, c_i32p, c_f32p, c_f32p, c_f32p, c_f32p, c_u32p, c_u32p, c_u32p]),
    "relinearise_factors": (C.c_int, [C.c_void_p]),
    "prep_messages": (C.c_int, [C.c_void_p]),
    "compute_messages": (C.c_int, [C.c_void_p]),
    "update_beliefs": (C.c_int, [C.c_void_p]),
    "weaken_prior_vertices": (C.c_int, [C.c_void_p]),
    "tensor_nbytes": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t)]),
    "get_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "dims": (C.c_int, [C.c_void_p, c_u32p, c_u32p, c_u32p, c_u32p, c_u32p]),
}

# Entry points only libgbp_cuda.so has.
CUDA_ONLY_API = {
    "gbp_cuda_last_error": (C.c_char_p, []),
    "gbp_cuda_version": (C.c_char_p, []),
    "gbp_opts_default": (None, [C.POINTER(GbpOpts)]),
    "gbp_cuda_last_timing": (C.c_int, [C.c_void_p, c_f32p, C.POINTER(C.c_uint64)]),
    "gbp_cuda_iterate_until": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(GbpIterStats),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gbp_cuda_add_keyframe_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]),
    "gbp_cuda_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "gbp_cuda_last_kernel_times": (C.c_int, [C.c_void_p, c_f32p, c_f32p]),
    "gbp_cuda_last_sweep_times": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int, C.POINTER(C.c_int)]),
    "gbp_cuda_debug_timestamps": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]),
    "gbp_cuda_iterate_async": (C.c_int, [C.c_void_p, C.c_int]),
    "gbp_cuda_synchronize": (C.c_int, [C.c_void_p]),
    "gbp_cuda_stream": (C.c_void_p, [C.c_void_p]),
    "gbp_cuda_plan_shard": (C.c_int, [C.POINTER(GbpProblem), C.c_uint32, C.c_uint32, C.POINTER(GbpShardPlan)]),
    "gbp_cuda_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "gbp_cuda_init_shard": (C.c_int, [C.POINTER(GbpProblem), C.POINTER(GbpOpts), C.c_uint32, C.c_uint32,
                                      C.c_void_p, C.POINTER(C.c_void_p)]),
    "gbp_cuda_init_group": (C.c_int, [C.POINTER(GbpProblem), C.POINTER(GbpOpts), C.c_uint32, C.POINTER(C.c_int),
                                      C.POINTER(C.c_void_p)]),
    "gbp_cuda_group_iterate": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.c_int, C.POINTER(GbpIterStats)]),
    "gbp_cuda_group_weaken_priors": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32]),
    "gbp_cuda_group_eval": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(GbpIterStats)]),
    "gbp_cuda_group_free": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32]),
    "gbp_cuda_release_cached_memory": (C.c_int, []),
    "gbp_cuda_test_inv6x6": (C.c_int, [c_f32p, c_f32p, C.c_int]),
    "gbp_cuda_test_inv3x3": (C.c_int, [c_f32p, c_f32p, C.c_int]),
    "gbp_cuda_test_project": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int]),
    "gbp_cuda_shard_info": (C.c_void_p, [C.c_void_p]),
    "gbp_cuda_exchange_mode": (C.c_int, [C.c_void_p]),
    # pure host: the rank-local sub-problem of a camera-range partition
    "gbp_shard_build": (C.c_int, [C.POINTER(GbpProblem), C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "gbp_shard_build_view": (C.c_int, [C.POINTER(GbpProblem), C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "gbp_shard_free": (None, [C.c_void_p]),
    "gbp_shard_problem": (C.POINTER(GbpProblem), [C.c_void_p]),
    "gbp_shard_get_plan": (C.POINTER(GbpShardPlan), [C.c_void_p]),
    "gbp_shard_lmk_global": (c_u32p, [C.c_void_p]),
    "gbp_shard_edge_global": (c_u32p, [C.c_void_p]),
    "gbp_shard_n_boundary_local": (C.c_uint32, [C.c_void_p]),
    "gbp_shard_boundary_local": (c_u32p, [C.c_void_p]),
    "gbp_shard_boundary_slot": (c_u32p, [C.c_void_p]),
    "gbp_shard_boundary_ranks": (c_u32p, [C.c_void_p]),
    "gbp_shard_n_active_global": (C.c_uint32, [C.c_void_p]),
    "gbp_shard_cam_bounds": (c_u32p, [C.c_void_p]),
}

HOST_API = {
    "gbp_cli_options_default": (None, [C.POINTER(GbpCliOptions)]),
    "gbp_bal_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "gbp_bal_from_arrays": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, c_f64p, c_u32p, c_u32p, c_f64p, c_f64p,
                                      c_f64p, C.POINTER(C.c_void_p)]),
    "gbp_bal_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "gbp_bal_with_means": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p, C.POINTER(C.c_void_p)]),
    "gbp_bal_free": (None, [C.c_void_p]),
    "gbp_bal_dims": (C.c_int, [C.c_void_p, c_u32p, c_u32p, c_u32p]),
    "gbp_bal_camera_index": (c_u32p, [C.c_void_p]),
    "gbp_bal_point_index": (c_u32p, [C.c_void_p]),
    "gbp_bal_observations": (c_f64p, [C.c_void_p]),
    "gbp_bal_parameters": (c_f64p, [C.c_void_p]),
    "gbp_bal_intrinsics": (c_f64p, [C.c_void_p]),
    "gbp_setup_create": (C.c_int, [C.c_void_p, C.POINTER(GbpCliOptions), C.c_int, C.POINTER(C.c_void_p)]),
    "gbp_setup_free": (None, [C.c_void_p]),
    "gbp_setup_problem": (C.POINTER(GbpProblem), [C.c_void_p]),
    "gbp_setup_next_keyframe": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_i32p,
                                          C.POINTER(C.c_int)]),
    "gbp_setup_data_counter": (C.c_int, [C.c_void_p]),
    "gbp_synth_generate": (C.c_int, [C.c_uint32, C.c_uint32, C.c_double, C.c_uint32, C.POINTER(C.c_void_p)]),
}


def declared_symbols():
    """Every symbol include/*.h declares (used by the symbol-export test)."""
    names = ["gbp_cuda_" + k for k in ENGINE_API]
    names += list(CUDA_ONLY_API) + list(HOST_API)
    return names


def bind_engine(lib, prefix):
    """Return {short name: ctypes function} for the engine API of `lib`."""
    out = {}
    for name, (res, args) in ENGINE_API.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args
        out[name] = fn
    return out


def bind_table(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


_lib = None


def load_library():
    """Load libgbp_cuda.so (built in-tree by __graft_entry__.build / csrc/Makefile).

    Fails loudly when the extension has not been built: there is no Python or
    CPU fallback for the hot path.
    """
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C gbp_poplar_b200/csrc`). There is no CPU fallback for the GBP hot path.")
        lib = C.CDLL(LIB_PATH)
        bind_table(lib, CUDA_ONLY_API)
        bind_table(lib, HOST_API)
        _lib = lib
    return _lib
