"""ctypes binding of libcadm_b200.so (the C ABI declared in include/cadm_b200.h).

There is no CPU fallback: if the shared library is missing, or there is no CUDA device, the product
raises.  Nothing here imports `oracle/`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcadm_b200.so")

# must mirror struct CadmConfig in include/cadm_b200.h
class CadmConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("env_id", C.c_int32), ("obs_dim", C.c_int32), ("proc_obs_dim", C.c_int32),
        ("act_dim", C.c_int32), ("ctx_dim", C.c_int32), ("hist_len", C.c_int32), ("hidden", C.c_int32),
        ("n_hidden", C.c_int32), ("enc_hidden", C.c_int32 * 3), ("ensemble", C.c_int32), ("particles", C.c_int32),
        ("candidates", C.c_int32), ("horizon", C.c_int32), ("m_max", C.c_int32), ("deterministic", C.c_int32),
        ("discrete", C.c_int32), ("num_elites", C.c_int32), ("cem_iters", C.c_int32), ("alpha", C.c_float),
        ("precision", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("context_layout", C.c_int32),
        ("max_torque", C.c_float),
    ]


# must mirror struct CadmTrainConfig in include/cadm_b200.h
class CadmTrainConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("env_id", C.c_int32), ("obs_dim", C.c_int32), ("proc_obs_dim", C.c_int32),
        ("act_dim", C.c_int32), ("ctx_dim", C.c_int32), ("hist_len", C.c_int32), ("hidden", C.c_int32),
        ("n_hidden", C.c_int32), ("enc_hidden", C.c_int32 * 3), ("ensemble", C.c_int32), ("deterministic", C.c_int32),
        ("has_back", C.c_int32), ("back_coeff", C.c_float), ("weight_decay_coeff", C.c_float), ("learning_rate", C.c_float),
        ("weight_decays", C.c_float * 8), ("context_weight_decays", C.c_float * 4),
        ("adam_beta1", C.c_float), ("adam_beta2", C.c_float), ("adam_eps", C.c_float),
    ]


ENV_IDS = {"halfcheetah": 0, "cripple_halfcheetah": 0, "ant": 1, "slim_humanoid": 2, "cartpole": 3, "pendulum": 4}
PRECISIONS = {"fp32": 0, "tc3x": 1, "tc1x": 2}
CTX_LAYOUTS = {"reference": 0, "matched": 1}

_P = C.c_void_p
_F = C.c_void_p      # device float* (passed as integer addresses)

# every symbol include/cadm_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cadm_abi_version": (C.c_int, []),
    "cadm_last_error": (C.c_char_p, [_P]),
    "cadm_plan_create": (C.c_int, [C.POINTER(CadmConfig), C.POINTER(_P)]),
    "cadm_plan_destroy": (C.c_int, [_P]),
    "cadm_plan_set_weights": (C.c_int, [_P, C.POINTER(_F), C.POINTER(_F), C.c_int32, _F, _F, _P]),
    "cadm_plan_set_encoder": (C.c_int, [_P, C.POINTER(_F), C.POINTER(_F), C.c_int32, _P]),
    "cadm_plan_set_norm": (C.c_int, [_P] + [_F] * 10 + [_P]),
    "cadm_encode_context": (C.c_int, [_P, C.c_int32, _F, _F, _F, _P]),
    "cadm_predict": (C.c_int, [_P, C.c_int32, _F, _F, _F, _F, C.c_uint64, _F, _F, _F, _P]),
    "cadm_rollout": (C.c_int, [_P, C.c_int32, C.c_int32, _F, _F, _F, _F, C.c_uint64, _F, _F, _P]),
    "cadm_cem_begin": (C.c_int, [_P, C.c_int32, _F, _F, _F, _F, _F, _P]),
    "cadm_cem_rollout": (C.c_int, [_P, C.c_int32, C.c_uint64, _F, _F, _P]),
    "cadm_cem_returns_buffer": (C.c_void_p, [_P]),
    "cadm_cem_returns_slice_elems": (C.c_int64, [_P]),
    "cadm_peer_export": (C.c_int, [_P, C.c_void_p]),
    "cadm_peer_attach": (C.c_int, [_P, C.c_void_p, C.c_int32]),
    "cadm_peer_enabled": (C.c_int, [_P]),
    "cadm_cem_refit": (C.c_int, [_P, C.c_int32, _P]),
    "cadm_cem_finish": (C.c_int, [_P, _F, _F, _F, _F, _P]),
    "cadm_plan_cem": (C.c_int, [_P, C.c_int32, _F, _F, _F, _F, _F, C.c_uint64, _F, _F, _F, _F, _F, _F, _P]),
    "cadm_plan_cem_host": (C.c_int, [_P, C.c_int32, _F, _F, _F, _F, _F, C.c_uint64, _F, _P]),
    "cadm_session_reset": (C.c_int, [_P, C.c_int32, C.c_void_p, _P]),
    "cadm_session_act": (C.c_int, [_P, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p, _P]),
    "cadm_session_observe": (C.c_int, [_P, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, _P]),
    "cadm_session_state": (C.c_int, [_P, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _P]),
    "cadm_plan_rs": (C.c_int, [_P, C.c_int32, _F, _F, _F, C.c_uint64, _F, _F, _F, _F, _F, _F, _F, _P]),
    "cadm_set_precision": (C.c_int, [_P, C.c_int32]),
    "cadm_selftest_tc_gemm": (C.c_int, [_F, _F, C.c_int32, C.c_int32, C.c_int32, _F, _P]),
    "cadm_selftest_tcs_gemm": (C.c_int, [_F, _F, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _F, _P]),
    "cadm_set_option": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "cadm_selftest_tc_rate": (C.c_int, [C.c_int32] * 6 + [C.c_void_p]),
    "cadm_selftest_tcs_rate": (C.c_int, [C.c_int32] * 5 + [C.c_void_p]),
    "cadm_debug_trace": (C.c_int, [_P, C.c_void_p, C.c_int32]),
    "cadm_launch_count": (C.c_int64, [_P]),
    "cadm_kernel_name": (C.c_char_p, [_P]),
    "cadm_set_timing": (C.c_int, [_P, C.c_int32]),
    "cadm_last_rollout_ms": (C.c_float, [_P]),
    # fit() on the device (csrc/trainer.cu)
    "cadm_train_create": (C.c_int, [C.POINTER(CadmTrainConfig), C.POINTER(_P)]),
    "cadm_train_destroy": (C.c_int, [_P]),
    "cadm_train_last_error": (C.c_char_p, [_P]),
    "cadm_train_param_count": (C.c_int64, [_P]),
    "cadm_train_launch_count": (C.c_int64, [_P]),
    "cadm_train_set_params": (C.c_int, [_P, C.c_void_p, C.c_int64]),
    "cadm_train_get_params": (C.c_int, [_P, C.c_void_p, C.c_int64]),
    "cadm_train_get_grads": (C.c_int, [_P, C.c_void_p, C.c_int64]),
    "cadm_train_adam_state": (C.c_int, [_P, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "cadm_train_set_norm": (C.c_int, [_P, C.POINTER(C.c_void_p), C.c_int32]),
    "cadm_train_set_dataset": (C.c_int, [_P, C.c_int32, C.c_int64] + [C.c_void_p] * 7),
    "cadm_train_step": (C.c_int, [_P, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
}

_lib = None


class CadmError(RuntimeError):
    pass


def load():
    """Load the library (once) and type every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CadmError(
            f"{LIB_PATH} is missing: build it with `python -m cadm_b200.build` "
            "(cadm_b200 has no CPU fallback; the CUDA extension is required)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.cadm_abi_version() != 1:
        raise CadmError("libcadm_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(handle, code):
    if code != 0:
        msg = load().cadm_last_error(handle)
        raise CadmError(f"cadm_b200 error {code}: {msg.decode() if msg else '?'}")
