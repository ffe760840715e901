"""ctypes binding of libdmfg.so (include/dmfg.h).

The library is built in-tree (``python -m discrete_mean_field_game_b200.build``
or ``__graft_entry__.build()``) and loaded from the package directory.  There
is deliberately NO fallback: if the shared object is missing or a call fails,
the product raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMFG_LIB_PATH") or os.path.join(_HERE, "libdmfg.so")   # override: A/B builds of the kernels

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4
F32, F64 = 0, 1
REWARD_NONE, REWARD_AC2, REWARD_SYNTHETIC = 0, 1, 2
DISCOUNT_STEP, DISCOUNT_CUMULATIVE = 0, 1
NOISE_INJECTED, NOISE_PHILOX, NOISE_ACTIONS = 0, 1, 2
VARIANT_AUTO, VARIANT_GENERIC, VARIANT_FAST, VARIANT_V2 = 0, 1, 2, 3
DROPOUT_NONE, DROPOUT_MASKS, DROPOUT_PHILOX = 0, 1, 2
MAX_D = 256

REWARD_KINDS = {"none": REWARD_NONE, "ac2": REWARD_AC2, "synthetic": REWARD_SYNTHETIC}
VARIANTS = {"auto": VARIANT_AUTO, "generic": VARIANT_GENERIC, "fast": VARIANT_FAST, "v2": VARIANT_V2}


class DmfgError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libdmfg error %d: %s" % (code, message))
        self.code = code


class RolloutArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("d", C.c_int32), ("T", C.c_int32),
        ("B", C.c_int64), ("pop_offset", C.c_int64),
        ("theta", C.c_double), ("theta_dev", C.c_void_p),
        ("shift", C.c_double), ("alpha_scale", C.c_double), ("gamma", C.c_double),
        ("reward_kind", C.c_int32), ("discount_kind", C.c_int32),
        ("noise_kind", C.c_int32), ("variant", C.c_int32),
        ("noise_y", C.c_void_p), ("seed", C.c_uint64), ("step_offset", C.c_uint64),
        ("pi0", C.c_void_p), ("w", C.c_void_p), ("rewards_in", C.c_void_p),
        ("states", C.c_void_p), ("actions", C.c_void_p), ("alpha", C.c_void_p),
        ("alpha_deriv", C.c_void_p), ("rewards", C.c_void_p), ("deltas", C.c_void_p),
        ("grads", C.c_void_p), ("pi_final", C.c_void_p), ("acc", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64), ("step_offset_dev", C.c_void_p),
    ]


class TdArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("d", C.c_int32), ("T", C.c_int32),
        ("B", C.c_int64), ("gamma", C.c_double), ("discount_kind", C.c_int32), ("reserved", C.c_int32),
        ("states", C.c_void_p), ("rewards", C.c_void_p), ("grads", C.c_void_p), ("w", C.c_void_p),
        ("deltas", C.c_void_p), ("acc", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
    ]


class LearnersArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("d", C.c_int32), ("T", C.c_int32),
        ("L", C.c_int64), ("learner_offset", C.c_int64), ("E", C.c_int32), ("episode0", C.c_int32),
        ("theta", C.c_void_p), ("w", C.c_void_p), ("shift", C.c_void_p), ("alpha_scale", C.c_void_p),
        ("shift_scalar", C.c_double), ("alpha_scale_scalar", C.c_double), ("gamma", C.c_double),
        ("lr_critic", C.c_double), ("lr_actor", C.c_double),
        ("constant_lr", C.c_int32), ("reward_kind", C.c_int32), ("discount_kind", C.c_int32),
        ("noise_kind", C.c_int32),
        ("mat_pi0", C.c_void_p), ("S", C.c_int32), ("layout", C.c_int32),
        ("start_rows", C.c_void_p), ("noise_y", C.c_void_p), ("seed", C.c_uint64),
        ("noise_episode_offset", C.c_int64),
        ("theta_trace", C.c_void_p), ("delta_trace", C.c_void_p), ("total_reward", C.c_void_p),
        ("pi_final", C.c_void_p),
    ]


class RnetArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("d", C.c_int32), ("n_fc3", C.c_int32), ("n_fc4", C.c_int32),
        ("N", C.c_int64), ("params", C.c_void_p), ("states", C.c_void_p), ("actions", C.c_void_p),
        ("dropout", C.c_int32), ("keep_prob", C.c_float), ("mask3", C.c_void_p), ("mask4", C.c_void_p),
        ("seed", C.c_uint64), ("sample_offset", C.c_uint64), ("rewards", C.c_void_p),
        ("drewards", C.c_void_p), ("grad", C.c_void_p), ("accumulate", C.c_int32), ("reserved", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
        ("gather_T", C.c_int32), ("reserved2", C.c_int32), ("gather_slots", C.c_int32 * 32),
    ]


MAX_GATHER = 32


class IrlNetArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("n_fc3", C.c_int32), ("n_fc4", C.c_int32), ("dropout", C.c_int32),
        ("keep_prob", C.c_float), ("reserved", C.c_int32), ("params", C.c_void_p),
        ("seed", C.c_uint64), ("sample_offset", C.c_uint64), ("reward_trace", C.c_void_p),
    ]


class IrlLossArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("T", C.c_int32), ("n_demo", C.c_int64), ("M", C.c_int64),
        ("gen_t_stride", C.c_int64), ("gen_j_stride", C.c_int64), ("num_demo_traj", C.c_double),
        ("r_demo", C.c_void_p), ("r_gen", C.c_void_p), ("log_z", C.c_void_p),
        ("d_demo", C.c_void_p), ("d_gen", C.c_void_p), ("loss_out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
    ]


class IrlGenArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("T", C.c_int32), ("M", C.c_int64),
        ("gen_t_stride", C.c_int64), ("gen_j_stride", C.c_int64), ("n_demo", C.c_int64),
        ("num_demo_traj", C.c_double), ("r_demo", C.c_void_p), ("loss_out", C.c_void_p),
        ("local_sums", C.c_int32), ("reserved", C.c_int32),
    ]


class IrlStepArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("l1l2", C.c_int32), ("params", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
        ("step", C.c_int64), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
        ("reg_loss_out", C.c_void_p),
    ]


# every symbol include/dmfg.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("dmfg_version", C.c_int, []),
    ("dmfg_last_error", C.c_char_p, []),
    ("dmfg_kernel_launches", C.c_uint64, []),
    ("dmfg_num_features", C.c_int64, [C.c_int32]),
    ("dmfg_acc_len", C.c_int64, [C.c_int32]),
    ("dmfg_rollout_workspace_bytes", C.c_uint64, [C.POINTER(RolloutArgs)]),
    ("dmfg_rollout", C.c_int, [C.POINTER(RolloutArgs), C.c_void_p]),
    ("dmfg_rollout_host", C.c_int, [C.POINTER(RolloutArgs), C.c_void_p]),
    ("dmfg_td_workspace_bytes", C.c_uint64, [C.POINTER(TdArgs)]),
    ("dmfg_td_accumulate", C.c_int, [C.POINTER(TdArgs), C.c_void_p]),
    ("dmfg_critic_eval", C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    ("dmfg_traj_metrics", C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("dmfg_ac_apply_update", C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                       C.c_double, C.c_double, C.c_void_p]),
    ("dmfg_ac_apply_update_dev", C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                           C.c_void_p]),
    ("dmfg_ac_learners", C.c_int, [C.POINTER(LearnersArgs), C.c_void_p]),
    ("dmfg_irl_learners", C.c_int, [C.POINTER(LearnersArgs), C.POINTER(IrlNetArgs), C.c_void_p]),
    ("dmfg_ac_step_workspace_bytes", C.c_uint64, [C.POINTER(RolloutArgs)]),
    ("dmfg_ac_step", C.c_int, [C.POINTER(RolloutArgs), C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                               C.c_void_p, C.c_void_p]),
    ("dmfg_rnet_param_count", C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    ("dmfg_rnet_param_offsets", C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    ("dmfg_rnet_workspace_bytes", C.c_uint64, [C.POINTER(RnetArgs)]),
    ("dmfg_rnet_forward", C.c_int, [C.POINTER(RnetArgs), C.c_void_p]),
    ("dmfg_rnet_backward", C.c_int, [C.POINTER(RnetArgs), C.c_void_p]),
    ("dmfg_irl_loss_workspace_bytes", C.c_uint64, [C.c_int64]),
    ("dmfg_irl_loss_grad", C.c_int, [C.POINTER(IrlLossArgs), C.c_void_p]),
    ("dmfg_rnet_backward_gen", C.c_int, [C.POINTER(RnetArgs), C.POINTER(IrlGenArgs), C.c_void_p]),
    ("dmfg_irl_reward_step_workspace_bytes", C.c_uint64, [C.POINTER(RnetArgs)]),
    ("dmfg_irl_reward_step", C.c_int, [C.POINTER(RnetArgs), C.POINTER(RnetArgs), C.POINTER(IrlGenArgs),
                                       C.POINTER(IrlStepArgs), C.c_void_p]),
    ("dmfg_irl_dp_finalize", C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("dmfg_adam_tf", C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int64,
                               C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, C.c_void_p, C.c_void_p]),
    ("dmfg_dirichlet_logq", C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_double, C.c_void_p, C.c_void_p]),
    ("dmfg_irl_log_z", C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_double,
                                 C.c_void_p, C.c_void_p]),
    ("dmfg_synthetic_check", C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    ("dmfg_umma_probe", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32,
                                  C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    ("dmfg_umma_selftest", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    ("dmfg_philox4x32_10", None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("dmfg_gamma_philox_rounds", C.c_int32, []),
    ("dmfg_philox4x32_gamma", None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("dmfg_gamma_sample", C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    ("dmfg_digamma", C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
]

_lib = None


def load():
    """Load libdmfg.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -m discrete_mean_field_game_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise DmfgError(rc, load().dmfg_last_error().decode("utf-8", "replace"))


def num_features(d):
    return d * (d + 1) // 2 + d + 1
