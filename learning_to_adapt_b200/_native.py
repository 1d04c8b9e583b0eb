"""ctypes binding of libl2a_b200.so (include/l2a_b200.h).  There is no CPU fallback: if the shared library is
missing or no B200 is present the product path raises."""
import ctypes as C
import os

from .build import DEBUG_LIB_PATH, LIB_PATH

L2A_MAX_LAYERS = 8

REWARD_HALF_CHEETAH, REWARD_ANT, REWARD_ARM = 0, 1, 2
SETS_SHARED, SETS_PER_ENV, SETS_ENSEMBLE_MEAN = 0, 1, 2
KERNEL_AUTO, KERNEL_SIMT, KERNEL_TCGEN05, KERNEL_TCGEN05_PAIR = 0, 1, 2, 3

EXPORTS = [
    "l2a_last_error", "l2a_version", "l2a_ctx_create", "l2a_ctx_destroy", "l2a_ctx_launch_count",
    "l2a_model_create", "l2a_model_destroy", "l2a_model_set_params", "l2a_model_get_params",
    "l2a_model_set_normalization", "l2a_model_param_block", "l2a_model_refresh", "l2a_rollout", "l2a_predict", "l2a_adapt", "l2a_cem_sample", "l2a_cem_refit",
    "l2a_shard_pack", "l2a_shard_select", "l2a_rnn_model_create", "l2a_rnn_model_destroy",
    "l2a_rnn_model_set_params", "l2a_rnn_model_set_normalization", "l2a_rnn_rollout", "l2a_rnn_predict",
    "l2a_window_create", "l2a_window_destroy", "l2a_window_set_normalization", "l2a_window_push", "l2a_window_reset",
    "l2a_window_length", "l2a_window_gather", "l2a_adapt_from_window", "l2a_plan_attach_window",
    "l2a_plan_create", "l2a_plan_run", "l2a_plan_destroy", "l2a_plan_create_ex", "l2a_plan_run_ex", "l2a_plan_exchange_buffer",
    "l2a_plan_attach_peers", "l2a_plan_exchange_resident", "l2a_ipc_get_handle", "l2a_ipc_open_handle", "l2a_ipc_close_handle", "l2a_plan_uses_graph", "l2a_plan_copy_candidates", "l2a_plan_copy_returns", "l2a_plan_io_bytes", "l2a_sample_uniform", "l2a_tc_plan_query", "l2a_tc2_plan_query",
]


# only in the debug build (lib/libl2a_b200_debug.so, -DL2A_DEBUG_KERNELS)
DEBUG_EXPORTS = ["l2a_debug_umma_tile", "l2a_debug_stream", "l2a_debug_set_timeline", "l2a_debug_mma_rate", "l2a_debug_pair"]


class MlpDesc(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("n_hidden", C.c_int32),
                ("hidden", C.c_int32 * (L2A_MAX_LAYERS - 1)), ("n_sets", C.c_int32)]


class RolloutParams(C.Structure):
    _fields_ = [("n_candidates", C.c_int32), ("n_envs", C.c_int32), ("horizon", C.c_int32), ("set_mode", C.c_int32),
                ("first_set", C.c_int32), ("n_sets", C.c_int32), ("reward_kind", C.c_int32), ("dt", C.c_float),
                ("act_stride_t", C.c_int64), ("act_stride_row", C.c_int64), ("kernel", C.c_int32),
                ("reserved", C.c_int32)]


SAMPLER_PHILOX, SAMPLER_MT19937 = 0, 1
PLANNER_RS, PLANNER_CEM = 0, 1
PLAN_ADAPT, PLAN_PUSH = 1, 2


class PlanOpts(C.Structure):
    _fields_ = [("sampler", C.c_int32), ("shard_rank", C.c_int32), ("shard_world", C.c_int32), ("n_candidates_total", C.c_int32),
                ("shard_offset", C.c_int64), ("seed", C.c_uint64), ("planner", C.c_int32), ("cem_iters", C.c_int32),
                ("cem_num_elites", C.c_int32), ("cem_compat", C.c_int32), ("cem_alpha", C.c_double)]


class PlanIO(C.Structure):
    _fields_ = [("mt_key", C.c_void_p), ("mt_pos", C.c_void_p), ("act_out", C.c_void_p), ("ret_out", C.c_void_p),
                ("idx_out", C.c_void_p), ("mt_has_gauss", C.c_void_p), ("mt_cached", C.c_void_p), ("cem_mean_out", C.c_void_p),
                ("cem_std_out", C.c_void_p), ("flags", C.c_int32), ("reserved", C.c_int32)]


_libs = {}


def load(debug=False):
    """Load the shared library (once per flavour) and declare the prototypes.  Raises ImportError when it is missing.
    debug=True: the debug build, which also exports the l2a_debug_* diagnostics (tests of the MMA tile path, probes)."""
    if debug in _libs:
        return _libs[debug]
    default = DEBUG_LIB_PATH if debug else LIB_PATH
    path = os.environ.get("L2A_B200_DEBUG_LIB" if debug else "L2A_B200_LIB", default)      # development: A/B two builds on the same box
    if not os.path.exists(path):
        raise ImportError(
            "learning_to_adapt_b200: %s is missing. Build it with `python -m learning_to_adapt_b200.build%s` "
            "(nvcc, sm_100a). There is no CPU fallback for the planning path." % (path, " --debug" if debug else ""))
    lib = C.CDLL(path)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    pp = C.POINTER(vp)
    lib.l2a_last_error.restype = C.c_char_p
    lib.l2a_last_error.argtypes = []
    lib.l2a_version.restype = i32
    lib.l2a_ctx_create.argtypes = [i32, pp]
    lib.l2a_ctx_destroy.argtypes = [vp]
    lib.l2a_ctx_launch_count.argtypes = [vp]
    lib.l2a_ctx_launch_count.restype = i64
    lib.l2a_model_create.argtypes = [vp, C.POINTER(MlpDesc), pp]
    lib.l2a_model_destroy.argtypes = [vp, vp]
    lib.l2a_model_set_params.argtypes = [vp, vp, i32, pp, pp, vp]
    lib.l2a_model_get_params.argtypes = [vp, vp, i32, pp, pp, vp]
    lib.l2a_model_set_normalization.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.l2a_model_param_block.argtypes = [vp, vp, i32, pp, C.POINTER(C.c_int64), vp, vp]
    lib.l2a_model_refresh.argtypes = [vp, vp, i32, i32, vp]
    lib.l2a_rollout.argtypes = [vp, vp, C.POINTER(RolloutParams), vp, vp, vp, vp, vp, vp, vp, vp]
    lib.l2a_plan_create.argtypes = [vp, vp, C.POINTER(RolloutParams), f32, vp, vp, C.c_uint64, pp]
    lib.l2a_plan_run.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.l2a_plan_destroy.argtypes = [vp, vp]
    lib.l2a_plan_create_ex.argtypes = [vp, vp, C.POINTER(RolloutParams), f64, vp, vp, C.POINTER(PlanOpts), pp]
    lib.l2a_plan_run_ex.argtypes = [vp, vp, vp, C.POINTER(PlanIO), vp]
    lib.l2a_plan_exchange_buffer.argtypes = [vp, pp, C.POINTER(C.c_uint64)]
    lib.l2a_plan_attach_peers.argtypes = [vp, vp, vp]
    lib.l2a_plan_exchange_resident.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.l2a_ipc_get_handle.argtypes = [vp, vp, vp]
    lib.l2a_ipc_open_handle.argtypes = [vp, vp, pp]
    lib.l2a_ipc_close_handle.argtypes = [vp, vp]
    lib.l2a_plan_uses_graph.argtypes = [vp]
    lib.l2a_plan_copy_candidates.argtypes = [vp, vp, vp]
    lib.l2a_plan_copy_returns.argtypes = [vp, vp, vp]
    lib.l2a_plan_io_bytes.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.l2a_tc_plan_query.argtypes = [C.POINTER(MlpDesc), vp]
    lib.l2a_tc2_plan_query.argtypes = [C.POINTER(MlpDesc), i32, i32, i32, i32, vp]
    lib.l2a_sample_uniform.argtypes = [vp, vp, vp, vp, i64, i32, C.c_uint64, C.c_uint64, vp]
    lib.l2a_predict.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, vp, vp, i32, vp]
    lib.l2a_adapt.argtypes = [vp, vp, vp, vp, i32, i32, f32, i32, i32, vp]
    lib.l2a_window_create.argtypes = [vp, i32, i32, i32, i32, pp]
    lib.l2a_window_destroy.argtypes = [vp, vp]
    lib.l2a_window_set_normalization.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.l2a_window_push.argtypes = [vp, vp, vp, vp, vp]
    lib.l2a_window_reset.argtypes = [vp, vp, i32, vp]
    lib.l2a_window_length.argtypes = [vp, vp, i32]
    lib.l2a_window_gather.argtypes = [vp, vp, vp, vp, vp]
    lib.l2a_adapt_from_window.argtypes = [vp, vp, vp, f32, i32, i32, vp]
    lib.l2a_plan_attach_window.argtypes = [vp, vp, vp, f32, i32, i32]
    lib.l2a_cem_sample.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.l2a_cem_refit.argtypes = [vp, vp, vp, i32, i32, i32, i32, f64, i32, vp, vp, vp, vp]
    lib.l2a_rnn_model_create.argtypes = [vp, i32, i32, i32, pp]
    lib.l2a_rnn_model_destroy.argtypes = [vp, vp]
    lib.l2a_rnn_model_set_params.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.l2a_rnn_model_set_normalization.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.l2a_rnn_rollout.argtypes = [vp, vp, C.POINTER(RolloutParams), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.l2a_rnn_predict.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.l2a_shard_pack.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp, vp]
    lib.l2a_shard_select.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    if debug:
        lib.l2a_debug_umma_tile.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
        lib.l2a_debug_set_timeline.argtypes = [vp, vp]
        lib.l2a_debug_mma_rate.argtypes = [vp, i32, i32, i32, vp, vp]
        lib.l2a_debug_pair.argtypes = [vp, i32, i32, i32, vp, vp]
        lib.l2a_debug_stream.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    for name in EXPORTS + (DEBUG_EXPORTS if debug else []):
        fn = getattr(lib, name)
        if name not in ("l2a_last_error", "l2a_ctx_launch_count"):
            fn.restype = i32
    _libs[debug] = lib
    return lib


def check(status):
    if status != 0:
        msgs = [lib.l2a_last_error().decode() for lib in _libs.values()]
        raise RuntimeError("libl2a_b200: %s (status %d)" % (" / ".join(m for m in msgs if m) or "error", status))
