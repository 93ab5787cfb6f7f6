"""ctypes binding of libtdmpc_b200.so (C ABI: include/tdmpc_b200.h).

There is deliberately no fallback: if the shared library is missing or fails to load, importing
the planner raises, and on a box without an sm_100 GPU `tdmpc_create` returns TDMPC_ERR_ARCH.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtdmpc_b200.so")
CSRC = os.path.join(_HERE, "csrc")

TDMPC_MAX_ITERS = 16
PREC_FP32, PREC_BF16, PREC_BF16_EXACT = 0, 1, 2
DYN_MLP, DYN_DRNN = 0, 1

EXPORTS = [
    "tdmpc_create", "tdmpc_destroy", "tdmpc_last_error", "tdmpc_reserve", "tdmpc_set_weights",
    "tdmpc_encode", "tdmpc_estimate_value", "tdmpc_estimate_value_drnn", "tdmpc_pi_rollout", "tdmpc_select_update", "tdmpc_plan",
    "tdmpc_fill_noise", "tdmpc_launch_count", "tdmpc_build_info", "tdmpc_abi_sizes",
    "tdmpc_replay_workspace_bytes", "tdmpc_replay_sample", "tdmpc_replay_gather", "tdmpc_update_forward",
    "tdmpc_train_param_count", "tdmpc_train_workspace_bytes", "tdmpc_update",
]

_fp = C.c_void_p  # device pointers travel as integers


class ModelCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "obs_dim", "action_dim", "latent_dim", "mlp_dim", "enc_dim", "hidden_dim",
        "enc_layernorm", "dynamics", "precision", "device")]


class Weights(C.Structure):
    _fields_ = [
        ("enc_w", _fp * 2), ("enc_b", _fp * 2), ("enc_ln_g", _fp), ("enc_ln_b", _fp),
        ("dyn_w", _fp * 3), ("dyn_b", _fp * 3), ("rew_w", _fp * 3), ("rew_b", _fp * 3),
        ("pi_w", _fp * 3), ("pi_b", _fp * 3),
        ("q_w", (_fp * 3) * 2), ("q_b", (_fp * 3) * 2), ("q_ln_g", (_fp * 2) * 2), ("q_ln_b", (_fp * 2) * 2),
        ("gru_w_ih", _fp), ("gru_w_hh", _fp), ("gru_ln_g", _fp * 3), ("gru_ln_b", _fp * 3),
        ("prior_w", _fp * 2), ("prior_b", _fp * 2), ("prior_bn", _fp * 4),
    ]


class PlanArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("horizon", C.c_int32), ("iterations", C.c_int32), ("num_elites", C.c_int32),
        ("num_pi_trajs", C.c_int32),
        ("n_samples", C.c_int32 * TDMPC_MAX_ITERS), ("n_reuse", C.c_int32 * TDMPC_MAX_ITERS),
        ("n_pi", C.c_int32 * TDMPC_MAX_ITERS),
        ("discount", C.c_double), ("momentum", C.c_double),
        ("temperature", C.c_float), ("min_std", C.c_float), ("std_floor", C.c_float), ("std_max", C.c_float),
        ("init_std", C.c_float),
        ("eval_mode", C.c_int32), ("t0_all", C.c_int32), ("warm_last", C.c_int32), ("best_a_last", C.c_int32),
        ("shift_elites", C.c_int32), ("has_prev", C.c_int32), ("elite_horizon", C.c_int32),
        ("color_len", C.c_int32), ("color_mats", _fp),
        ("n_white", C.c_int32 * TDMPC_MAX_ITERS), ("n_color1", C.c_int32 * TDMPC_MAX_ITERS),
        ("obs", _fp), ("z0", _fp), ("t0", _fp), ("prev_mean", _fp), ("elite_actions", _fp),
        ("hidden", _fp), ("hidden_out", _fp),
        ("seed", C.c_uint64), ("call", C.c_uint64), ("call_dev", _fp),
        ("noise_pi", _fp), ("noise_cem", _fp), ("noise_term", _fp), ("noise_out", _fp), ("noise_shift", _fp), ("u_pick", _fp),
        ("action", _fp), ("metrics", _fp),
        ("tap_values", _fp), ("tap_elite_idx", _fp), ("tap_mean", _fp), ("tap_std", _fp), ("tap_z0", _fp),
    ]


class ReplaySampleArgs(C.Structure):
    _fields_ = [
        ("priorities", _fp), ("total", C.c_int64), ("batch", C.c_int32), ("replace", C.c_int32),
        ("per_alpha", C.c_float), ("per_beta", C.c_float), ("uniforms", _fp), ("n_uniforms", C.c_int64),
        ("seed", C.c_uint64), ("call", C.c_uint64), ("max_rounds", C.c_int32), ("resume", C.c_int32),
        ("workspace", _fp), ("workspace_bytes", C.c_int64), ("idxs", _fp), ("weights", _fp), ("status", _fp),
    ]


class ReplayGatherArgs(C.Structure):
    _fields_ = [
        ("idxs", _fp), ("batch", C.c_int32), ("horizon", C.c_int32), ("obs_dim", C.c_int32), ("action_dim", C.c_int32),
        ("episode_length", C.c_int64), ("obs_buf", _fp), ("last_obs_buf", _fp), ("action_buf", _fp), ("reward_buf", _fp),
        ("obs", _fp), ("next_obs", _fp), ("action", _fp), ("reward", _fp),
    ]


class UpdateArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("horizon", C.c_int32), ("discount", C.c_double), ("rho", C.c_double), ("min_std", C.c_float),
        ("consistency_coef", C.c_float), ("reward_coef", C.c_float), ("value_coef", C.c_float),
        ("obs", _fp), ("next_obses", _fp), ("action", _fp), ("reward", _fp), ("noise_pi", _fp),
        ("seed", C.c_uint64), ("call", C.c_uint64),
        ("consistency_loss", _fp), ("reward_loss", _fp), ("value_loss", _fp), ("priority_loss", _fp), ("total_loss", _fp),
        ("zs", _fp), ("td_target", _fp), ("q1", _fp), ("q2", _fp), ("reward_pred", _fp),
    ]


class TrainArgs(C.Structure):
    """tdmpc_train_args (include/tdmpc_b200.h): the whole training step."""
    _fields_ = [
        ("online", Weights), ("target", Weights),
        ("batch", C.c_int32), ("horizon", C.c_int32), ("discount", C.c_double), ("rho", C.c_double), ("min_std", C.c_float),
        ("consistency_coef", C.c_float), ("reward_coef", C.c_float), ("value_coef", C.c_float),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("adam_eps", C.c_double),
        ("grad_clip_norm", C.c_double), ("tau", C.c_double), ("update_target", C.c_int32),
        ("adam_step", C.c_int64), ("pi_adam_step", C.c_int64),
        ("adam_m", _fp), ("adam_v", _fp), ("pi_adam_m", _fp), ("pi_adam_v", _fp), ("grads", _fp),
        ("obs", _fp), ("next_obses", _fp), ("action", _fp), ("reward", _fp), ("weights", _fp),
        ("noise_pi", _fp), ("noise_pi_update", _fp),
        ("seed", C.c_uint64), ("call", C.c_uint64),
        ("metrics", _fp), ("priorities", _fp), ("workspace", _fp), ("workspace_bytes", C.c_int64),
    ]


def build(verbose=False):
    """Compile libtdmpc_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8", "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:]); print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libtdmpc_b200.so failed")
    return LIB_PATH


_lib = None


def load():
    """Load the library (building it first only if the .so is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.tdmpc_create.argtypes = [C.POINTER(ModelCfg), C.POINTER(H)]
    lib.tdmpc_destroy.argtypes = [H]
    lib.tdmpc_last_error.argtypes = [H]; lib.tdmpc_last_error.restype = C.c_char_p
    lib.tdmpc_reserve.argtypes = [H, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.tdmpc_set_weights.argtypes = [H, C.POINTER(Weights), C.c_void_p]
    lib.tdmpc_encode.argtypes = [H, _fp, C.c_int32, _fp, C.c_void_p]
    lib.tdmpc_estimate_value.argtypes = [H, _fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_float,
                                         C.c_uint64, C.c_uint64, _fp, _fp, C.c_void_p]
    lib.tdmpc_estimate_value_drnn.argtypes = [H, _fp, _fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_float,
                                              C.c_uint64, C.c_uint64, _fp, _fp, C.c_void_p]
    lib.tdmpc_pi_rollout.argtypes = [H, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, C.c_uint64,
                                     _fp, C.c_void_p]
    lib.tdmpc_select_update.argtypes = [H, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_double,
                                        C.c_float, C.c_float, _fp, _fp, _fp, _fp, C.c_void_p]
    lib.tdmpc_plan.argtypes = [H, C.POINTER(PlanArgs), C.c_void_p]
    lib.tdmpc_fill_noise.argtypes = [H, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64, _fp, C.c_void_p]
    lib.tdmpc_launch_count.argtypes = [H]; lib.tdmpc_launch_count.restype = C.c_uint64
    lib.tdmpc_build_info.argtypes = []; lib.tdmpc_build_info.restype = C.c_char_p
    lib.tdmpc_abi_sizes.argtypes = [C.POINTER(C.c_int32), C.c_int32]; lib.tdmpc_abi_sizes.restype = C.c_int32
    lib.tdmpc_replay_workspace_bytes.argtypes = [C.c_int64, C.c_int32]; lib.tdmpc_replay_workspace_bytes.restype = C.c_int64
    lib.tdmpc_replay_sample.argtypes = [C.POINTER(ReplaySampleArgs), C.c_void_p]
    lib.tdmpc_replay_gather.argtypes = [C.POINTER(ReplayGatherArgs), C.c_void_p]
    lib.tdmpc_update_forward.argtypes = [H, H, C.POINTER(UpdateArgs), C.c_void_p]
    lib.tdmpc_train_param_count.argtypes = [C.POINTER(ModelCfg), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.tdmpc_train_param_count.restype = C.c_int64
    lib.tdmpc_train_workspace_bytes.argtypes = [C.POINTER(ModelCfg), C.c_int32, C.c_int32]
    lib.tdmpc_train_workspace_bytes.restype = C.c_int64
    lib.tdmpc_update.argtypes = [H, C.POINTER(TrainArgs), C.c_void_p]
    got = (C.c_int32 * 10)()
    lib.tdmpc_abi_sizes(got, 10)
    want = [C.sizeof(ModelCfg), C.sizeof(Weights), C.sizeof(PlanArgs), PlanArgs.obs.offset, PlanArgs.seed.offset,
            PlanArgs.action.offset, PlanArgs.tap_z0.offset, C.sizeof(ReplaySampleArgs), C.sizeof(ReplayGatherArgs), C.sizeof(UpdateArgs)]
    if list(got) != want:
        raise RuntimeError(f"libtdmpc_b200.so ABI mismatch: library {list(got)} vs binding {want} (stale build?)")
    for name in EXPORTS:
        if name not in ("tdmpc_last_error", "tdmpc_launch_count", "tdmpc_build_info", "tdmpc_abi_sizes", "tdmpc_replay_workspace_bytes"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib
