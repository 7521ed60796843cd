"""ctypes binding of libwdb200.so (the C ABI declared in include/wdb200.h).

This is the ONLY compute backend of the package: if the shared library is missing the
import of any kernel-facing module fails loudly -- there is no Python / CPU fallback.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwdb200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_ll = ctypes.c_longlong
_ull = ctypes.c_ulonglong


class ResetDesc(ctypes.Structure):
    """struct wdb_reset_desc (include/wdb200.h)."""

    _fields_ = [("dst", _vp), ("ref", _vp), ("bytes_per_env", _ll), ("pool_rows", _ll)]


_fp = ctypes.c_void_p  # device pointers travel as integers


class TcEnv(ctypes.Structure):
    """struct wdb_tc_env (include/wdb200.h)."""

    _fields_ = [
        ("n_envs", _i), ("n_agents", _i),
        ("loc_x", _fp), ("loc_y", _fp), ("speed", _fp), ("direction", _fp),
        ("acceleration", _fp), ("agent_types", _fp), ("edge_hit_reward_penalty", _fp),
        ("edge_hit_penalty", _f), ("grid_length", _f),
        ("acceleration_actions", _fp), ("turn_actions", _fp),
        ("max_speed", _f), ("num_other_agents_observed", _i), ("skill_levels", _fp),
        ("runner_exits_game_after_tagged", _i), ("still_in_the_game", _fp),
        ("use_full_observation", _i), ("obs", _fp), ("neighbor_distances", _fp),
        ("neighbor_ids_sorted_by_distance", _fp), ("nearest_neighbor_ids", _fp),
        ("rewards", _fp), ("step_rewards", _fp), ("num_runners", _fp),
        ("distance_margin_for_reward", _f), ("tag_reward_for_tagger", _f),
        ("tag_penalty_for_runner", _f), ("end_of_game_reward_for_runner", _f),
        ("done", _fp), ("env_timestep", _fp), ("episode_length", _i), ("stats", _fp),
        ("blocks_per_env", _i),
    ]


class TcPolicyIO(ctypes.Structure):
    """struct wdb_tc_policy_io."""

    _fields_ = [
        ("n_agents", _i), ("probs0", _fp), ("probs1", _fp), ("actions_batch", _fp),
        ("rewards_batch", _fp), ("obs_next", _fp), ("reward_running_sum", _fp),
        ("episodic_reward_sum", _fp), ("obs_next_tiles", _fp),
    ]


class TcRollout(ctypes.Structure):
    """struct wdb_tc_rollout."""

    _fields_ = [
        ("rng_state", _fp), ("uniforms", _fp),
        ("n_policies", _i), ("n_actions0", _i), ("n_actions1", _i),
        ("agent_policy", _fp), ("agent_slot", _fp),
        ("policy", TcPolicyIO * 4),
        ("sampled_actions", _fp), ("sampled_actions_0", _fp), ("sampled_actions_1", _fp),
        ("done_batch", _fp), ("step_running_sum", _fp),
        ("episodic_step_sum", _fp), ("num_completed_episodes", _fp),
        ("reset_table", _fp), ("n_reset_arrays", _i), ("obs_at_reset", _fp),
        ("reset_done_envs", _i), ("launch_after_forward", _i),
    ]


class GatherPolicy(ctypes.Structure):
    _fields_ = [("n_agents", _i), ("agent_ids", _fp), ("rows", _fp)]


class Gather(ctypes.Structure):
    """struct wdb_gather (include/wdb200.h)."""

    _fields_ = [
        ("n_envs", _i), ("n_agents", _i), ("width", _i), ("n_policies", _i), ("scatter", _i),
        ("full", _fp), ("policy", GatherPolicy * 4),
    ]


class BookkeepPolicy(ctypes.Structure):
    _fields_ = [
        ("n_agents", _i), ("agent_ids", _fp), ("rewards_batch", _fp), ("actions_batch", _fp),
        ("reward_running_sum", _fp), ("episodic_reward_sum", _fp),
    ]


class Bookkeep(ctypes.Structure):
    """struct wdb_bookkeep (include/wdb200.h)."""

    _fields_ = [
        ("n_envs", _i), ("n_agents", _i), ("n_policies", _i), ("n_heads", _i),
        ("done", _fp), ("rewards", _fp), ("actions", _fp), ("done_batch", _fp),
        ("step_running_sum", _fp), ("episodic_step_sum", _fp), ("num_completed_episodes", _fp),
        ("policy", BookkeepPolicy * 4),
    ]


class MlpPair(ctypes.Structure):
    """struct wdb_mlp_pair (include/wdb200.h)."""

    _fields_ = [
        ("blob", _fp * 2), ("F", _i * 2), ("H", _i * 2), ("A0", _i * 2), ("A1", _i * 2),
        ("obs", _fp * 2), ("rows", ctypes.c_longlong * 2),
        ("probs0", _fp * 2), ("probs1", _fp * 2), ("values", _fp * 2),
        ("ctas_b", _i), ("flags", _i),
    ]


MLP_WEIGHTS_STABLE = 1


class SaRollout(ctypes.Structure):
    """struct wdb_sa_rollout (include/wdb200.h)."""

    _fields_ = [
        ("env_kind", _i), ("n_envs", _i), ("n_steps", _i), ("episode_length", _i),
        ("state_dim", _i), ("use_argmax", _i), ("env_params", _f * 12),
        ("state", _fp), ("observations", _fp), ("done", _fp), ("env_timestep", _fp),
        ("rewards", _fp), ("sampled_actions", _fp),
        ("n_hidden", _i), ("dims", _i * 5), ("w", _fp * 4), ("b", _fp * 4),
        ("rng_state", _fp), ("uniforms", _fp),
        ("obs_batch", _fp), ("actions_batch", _fp), ("rewards_batch", _fp),
        ("done_batch", _fp), ("probs_batch", _fp),
        ("reward_running_sum", _fp), ("step_running_sum", _fp), ("episodic_reward_sum", _fp),
        ("episodic_step_sum", _fp), ("num_completed", _fp),
        ("reset_table", _fp), ("n_reset_arrays", _i), ("pool_rng", _fp),
        ("reset_done_envs", _i),
    ]


class PgLoss(ctypes.Structure):
    """struct wdb_pg_loss (include/wdb200.h)."""

    _fields_ = [
        ("T", _i), ("n_envs", _i), ("n_agents", _i), ("n_heads", _i), ("n_actions", _i * 4),
        ("probs", _fp * 4), ("values", _fp), ("actions", _fp), ("rewards", _fp), ("done", _fp),
        ("gamma", _f), ("vf_coeff", _f), ("entropy_coeff", _f),
        ("grad_probs", _fp * 4), ("grad_values", _fp), ("returns", _fp), ("sums", _fp),
    ]


_SIGNATURES = {
    "wdb_abi_version": (_i, []),
    "wdb_error_string": (ctypes.c_char_p, [_i]),
    "wdb_launch_count": (_ll, []),
    "wdb_set_option": (_i, [ctypes.c_char_p, _i]),
    "wdb_rng_state_bytes": (_ll, [_ll]),
    "wdb_rng_init": (_i, [_vp, _vp, _ll, _ull]),
    "wdb_rng_draw_u32x4": (_i, [_vp, _vp, _vp, _ll]),
    "wdb_sample_actions": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "wdb_sample_ou_process": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _vp]),
    "wdb_reset_when_done": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "wdb_reset_log_mask": (_i, [_vp, _vp, _i]),
    "wdb_update_log_mask": (_i, [_vp, _vp, _i, _i]),
    "wdb_log_one_step": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i]),
    "wdb_testkernel": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _f, _i, _i, _i]),
    "wdb_tag_gridworld_step": (
        _i,
        [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _i, _vp, _i, _vp],
    ),
    "wdb_tag_continuous_step": (
        _i,
        [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _f, _i,
         _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _vp,
         _vp, _i, _vp],
    ),
    "wdb_tag_continuous_rollout_step": (
        _i, [_vp, ctypes.POINTER(TcEnv), ctypes.POINTER(TcRollout)]),
    "wdb_mlp_blob_bytes": (_ll, [_i, _i, _i, _i]),
    "wdb_mlp_pack_weights": (_i, [_vp] * 12 + [_i, _i, _i, _i]),
    "wdb_mlp_policy_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _ll, _vp, _vp, _vp]),
    "wdb_mlp_policy_forward_pair": (_i, [_vp, ctypes.POINTER(MlpPair)]),
    "wdb_gather_policy_rows": (_i, [_vp, ctypes.POINTER(Gather)]),
    "wdb_rollout_bookkeep": (_i, [_vp, ctypes.POINTER(Bookkeep)]),
    "wdb_mlp_obs_tiles_bytes": (_ll, [_i, _ll]),
    "wdb_mlp_pack_obs": (_i, [_vp, _vp, _ll, _i, _vp]),
    "wdb_mlp_policy_forward_tiles": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _ll, _vp, _vp, _vp]),
    "wdb_discounted_returns": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f]),
    "wdb_pg_loss_and_grads": (_i, [_vp, ctypes.POINTER(PgLoss)]),
    "wdb_heads_softmax": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp, _vp, _vp]),
    "wdb_heads_softmax_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "wdb_pad_rows": (_i, [_vp, _vp, _ll, _i, _i, _vp]),
    "wdb_relu_backward_bias_rows": (_i, [_ll]),
    "wdb_relu_backward_bias": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "wdb_grad_sumsq": (_i, [_vp, _vp, _ll, _vp]),
    "wdb_adam_step": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _i, _f, _vp]),
    "wdb_single_agent_rollout_supported": (_i, [_i, ctypes.POINTER(_i)]),
    "wdb_single_agent_rollout": (_i, [_vp, ctypes.POINTER(SaRollout)]),
    "wdb_mountain_car_step": (
        _i, [_vp, _i, _vp, _vp, _vp, _vp, _vp] + [_f] * 7 + [_vp, _i]),
    "wdb_continuous_mountain_car_step": (
        _i, [_vp, _i, _vp, _vp, _vp, _vp, _vp] + [_f] * 8 + [_vp, _i]),
    "wdb_pendulum_step": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "wdb_acrobot_step": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "wdb_cartpole_step": (
        _i,
        [_vp, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _f, _f, _f, _vp, _i],
    ),
}

_lib = None


def exported_symbols():
    """Names every build must export (== the functions include/wdb200.h declares)."""
    return sorted(_SIGNATURES)


def load():
    """dlopen libwdb200.so and type its entry points.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m warp_drive_b200.build` "
            "(or __graft_entry__.build()).  warp_drive_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the build lost a symbol
        fn.restype = restype
        fn.argtypes = argtypes
    assert lib.wdb_abi_version() == 1
    _lib = lib
    # A/B switches for whole test / bench runs: WDB_OPTIONS="tc_variant=2,tc_v2_threads=224"
    for item in filter(None, os.environ.get("WDB_OPTIONS", "").split(",")):
        key, _, val = item.partition("=")
        rc = lib.wdb_set_option(key.strip().encode(), int(val))
        if rc != 0:
            raise RuntimeError(f"WDB_OPTIONS: wdb_set_option({key!r}, {val}) failed ({rc})")
    return lib


def check(err, what=""):
    if err != 0:
        msg = load().wdb_error_string(err).decode()
        raise RuntimeError(f"libwdb200 {what} failed: cudaError {err} ({msg})")


def ptr(t):
    """Device pointer of a torch CUDA tensor / DeviceArray handle (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "tensor"):
        t = t.tensor
    if torch.is_tensor(t):
        if not t.is_cuda:
            raise RuntimeError(
                "libwdb200 kernels need CUDA tensors; got a CPU tensor "
                "(warp_drive_b200 has no CPU fallback on the kernel path)"
            )
        if not t.is_contiguous():
            raise RuntimeError("libwdb200 kernels need C-contiguous tensors")
        return t.data_ptr()
    if hasattr(t, "gpudata"):  # pycuda-style pointer holder
        return int(t.gpudata)
    if hasattr(t, "data_ptr"):
        return int(t.data_ptr())
    if isinstance(t, int):
        return t
    raise TypeError(f"cannot take a device pointer of {type(t)}")


def stream_ptr(stream=None):
    """cudaStream_t of torch's current stream, so launches order with torch work."""
    if stream is None:
        stream = torch.cuda.current_stream()
    return stream.cuda_stream


def launch_count():
    return int(load().wdb_launch_count())
