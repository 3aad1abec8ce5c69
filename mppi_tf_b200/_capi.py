"""ctypes binding of include/mppi_b200.h (libmppi_b200.so).

There is no Python/CPU fallback here on purpose: if the shared library is missing, or no
sm_100 device is present, every compute entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPPI_B200_LIB: developer knob to load an experimental build of the same library
LIB_PATH = os.environ.get("MPPI_B200_LIB") or os.path.join(_HERE, "_build", "libmppi_b200.so")

MPPI_OK = 0
MPPI_ERR_BAD_ARG, MPPI_ERR_CUDA, MPPI_ERR_COMM, MPPI_ERR_UNSUPPORTED, MPPI_ERR_STATE = 1, 2, 3, 4, 5
MPPI_MAX_A = 8
MPPI_ACTION_COST_CPP, MPPI_ACTION_COST_PYTHON = 0, 1

_fp = C.POINTER(C.c_float)


class MppiConfig(C.Structure):
    """struct mppi_config, field for field."""
    _fields_ = [
        ("k", C.c_int), ("tau", C.c_int), ("s_dim", C.c_int), ("a_dim", C.c_int),
        ("dt", C.c_float), ("mass", C.c_float), ("lambda_", C.c_float),
        ("sigma", _fp), ("goal", _fp), ("q", _fp),
        ("seed", C.c_uint64),
        ("device", C.c_int), ("rank", C.c_int), ("world", C.c_int),
        ("n_controllers", C.c_int), ("goal_per_controller", C.c_int),
        ("stream", C.c_void_p),
        ("model", C.c_int),
        ("philox_rounds", C.c_int),
    ]


MODEL_POINT_MASS, MODEL_MLP, MODEL_AUV = 0, 1, 2


class MppiAuvParams(C.Structure):
    """struct mppi_auv_params, field for field."""
    _fields_ = [
        ("mass", C.c_float), ("volume", C.c_float), ("density", C.c_float),
        ("cog", C.c_float * 3), ("cob", C.c_float * 3),
        ("added_mass", C.c_float * 36), ("inertia", C.c_float * 6),
        ("linear_damping", C.c_float * 36), ("quad_damping", C.c_float * 6),
        ("linear_damping_forward_speed", C.c_float * 36),
        ("rk", C.c_int),
    ]


class MppiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mppi_b200 error {code}: {msg}")
        self.code = code


# every exported symbol of include/mppi_b200.h: name -> (restype, argtypes)
_i, _f, _u32, _u64, _vp = C.c_int, C.c_float, C.c_uint32, C.c_uint64, C.c_void_p
_H = C.c_void_p
SYMBOLS = {
    "mppi_version": (C.c_char_p, []),
    "mppi_config_default": (None, [C.POINTER(MppiConfig), _i, _i, _f, _f, _i, _i]),
    "mppi_create": (_i, [C.POINTER(MppiConfig), C.POINTER(_H)]),
    "mppi_destroy": (_i, [_H]),
    "mppi_last_error": (C.c_char_p, [_H]),
    "mppi_next": (_i, [_H, _fp, _fp]),
    "mppi_next_with_noise": (_i, [_H, _fp, _fp, _fp]),
    "mppi_next_with_noise_dev": (_i, [_H, _fp, _vp, _fp]),
    "mppi_set_state": (_i, [_H, _fp]),
    "mppi_enqueue_update": (_i, [_H, _vp]),
    "mppi_enqueue_exchange": (_i, [_H]),
    "mppi_enqueue_finish": (_i, [_H]),
    "mppi_fetch_action": (_i, [_H, _fp]),
    "mppi_synchronize": (_i, [_H]),
    "mppi_set_goal": (_i, [_H, _fp]),
    "mppi_set_goal_n": (_i, [_H, _fp, _i]),
    "mppi_set_lambda": (_i, [_H, _f]),
    "mppi_set_sigma": (_i, [_H, _fp]),
    "mppi_set_action_cost": (_i, [_H, _i, _f, _f]),
    "mppi_set_normalize_cost": (_i, [_H, _i]),
    "mppi_set_ellipse_cost": (_i, [_H] + [_f] * 7),
    "mppi_set_static_cost": (_i, [_H]),
    "mppi_cost_action_py": (_i, [_i, _i, _i, _f, _f, _f, _fp, _fp, _fp, _fp]),
    "mppi_cost_state_ellipse": (_i, [_i, _i, _fp] + [_f] * 7 + [_fp]),
    "mppi_set_q": (_i, [_H, _fp]),
    "mppi_set_action_limits": (_i, [_H, _i, _i, _fp, _fp]),
    "mppi_savgol_filter": (_i, [_i, _i, _fp, _i, _i, _fp]),
    "mppi_set_mass": (_i, [_H, _f]),
    "mppi_set_sequence": (_i, [_H, _fp]),
    "mppi_get_sequence": (_i, [_H, _fp]),
    "mppi_get_update": (_i, [_H, _fp]),
    "mppi_get_costs": (_i, [_H, _fp]),
    "mppi_get_weight_stats": (_i, [_H, _fp, _fp]),
    "mppi_set_update_counter": (_i, [_H, _u32]),
    "mppi_dump_noise": (_i, [_H, _fp]),
    "mppi_k_local": (_i, [_H]),
    "mppi_k_offset": (_i, [_H]),
    "mppi_exchange_stride": (_i, [_H]),
    "mppi_shard_range": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "mppi_payload_stride": (_i, [_i]),
    "mppi_exchange_buffers": (_i, [_H, C.POINTER(_vp), C.POINTER(_vp)]),
    "mppi_exchange_set_buffers": (_i, [_H, _vp, _vp]),
    "mppi_peer_handle": (_i, [_H, _vp]),
    "mppi_peer_attach": (_i, [_H, _vp]),
    "mppi_comm_unique_id": (_i, [_vp]),
    "mppi_comm_init": (_i, [_H, _vp]),
    "mppi_set_mlp": (_i, [_H, _i] + [_fp] * 10),
    "mppi_mlp_predict": (_i, [_H, _i, _i, _fp, _fp, _fp]),
    "mppi_mlp_train_step": (_i, [_H, _i, _fp, _fp, _fp, _f, _fp]),
    "mppi_mlp_set_adam": (_i, [_H, _f, _f, _f]),
    "mppi_mlp_train": (_i, [_H, _i, _fp, _fp, _fp, _i, _i, _f, _i, _f, _u64, _fp, C.POINTER(_i)]),
    "mppi_mlp_get_weights": (_i, [_H] + [_fp] * 6),
    "mppi_set_auv_model": (_i, [_H, C.POINTER(MppiAuvParams)]),
    "mppi_auv_predict": (_i, [_H, _i, _i, _fp, _fp, _fp]),
    "mppi_set_nn_auv_model": (_i, [_H, _i, _i, C.POINTER(_fp), C.POINTER(_fp), _fp, _fp, _fp, _fp]),
    "mppi_set_quat_cost": (_i, [_H, _fp]),
    "mppi_cost_state_quat": (_i, [_i, _i, _fp, _fp, _fp, _fp]),
    "mppi_set_ellipse3d_cost": (_i, [_H, _fp, _fp, _fp, _fp, _f, _f, _f]),
    "mppi_cost_state_ellipse3d": (_i, [_i, _i, _fp, _fp, _fp, _fp, _fp, _f, _f, _f, _fp]),
    "mppi_block_diag": (_i, [_fp, _i, _i, _i, _fp]),
    "mppi_model_free_step": (_i, [_i, _f, _f, _i, _i, _i, _fp, _fp]),
    "mppi_model_action_step": (_i, [_i, _f, _f, _i, _i, _i, _fp, _fp]),
    "mppi_model_step": (_i, [_i, _f, _f, _i, _i, _i, _i, _fp, _fp, _fp]),
    "mppi_cost_state": (_i, [_i, _i, _i, _fp, _fp, _fp, _fp]),
    "mppi_cost_action": (_i, [_i, _i, _i, _f, _fp, _fp, _fp, _fp]),
    "mppi_cost_step": (_i, [_i, _i, _i, _i, _f, _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "mppi_prepare_action": (_i, [_i, _i, _fp, _i, _fp]),
    "mppi_prepare_noise": (_i, [_i, _i, _i, _i, _fp, _i, _fp]),
    "mppi_update_stages": (_i, [_i, _i, _i, _i, _f, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "mppi_stage_vector_op": (_i, [_i, _i, _i, _fp, _f, _f, _fp]),
    "mppi_weighted_noise": (_i, [_i, _i, _i, _fp, _fp, _fp]),
    "mppi_get_new": (_i, [_i, _i, _fp, _i, _fp]),
    "mppi_shift": (_i, [_i, _i, _fp, _fp, _i, _fp]),
    "mppi_philox_raw": (_i, [_i, _u64, _u32, _u32, _u32, _u32, _i, C.POINTER(_u32)]),
    "mppi_debug_trace": (_i, [_H, _i]),
    "mppi_debug_get_trace": (_i, [_H, C.POINTER(C.c_uint64), _i]),
    "mppi_last_grid_x": (_i, [_H]),
    "mppi_philox_raw_rounds": (_i, [_i, _u64, _u32, _u32, _u32, _u32, _i, _i, C.POINTER(_u32)]),
}

_lib = None


def load():
    """Load libmppi_b200.so.  Raises (never falls back) when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C mppi_tf_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc != MPPI_OK:
        msg = load().mppi_last_error(handle)
        raise MppiError(rc, msg.decode() if msg else "unknown error")
