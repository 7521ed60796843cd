"""CPU oracle for the rollout hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under
``warp_drive_b200/`` imports it (tests/test_no_oracle_in_product.py enforces that).

Contents
  wd_oracle.c        C restatement of the reference *CUDA* kernels (the parity target)
  numpy_ref.py       NumPy restatement of the reference *Python CPU* env steps
                     (the reported CPU baseline)
  ref_cuda.py        loader for the reference's own CUDA kernels compiled to
                     oracle/_ref/*.fatbin (GPU-side bit-level oracle)
  build_ref.py       recipe that compiles those fatbins from /root/reference in place
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "wd_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libwdoracle.so")

_lib = None


def build(force=False):
    """gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC wd_oracle.c"""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if (
        not force
        and os.path.exists(_SO)
        and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)
    ):
        return _SO
    cmd = [
        "gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-Wall", "-shared", "-fPIC",
        _SRC, "-o", _SO, "-lm",
    ]
    subprocess.run(cmd, check=True)
    return _SO


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_c_int, _c_float, _c_long = ctypes.c_int, ctypes.c_float, ctypes.c_long


def _opt(ptr_type):
    """ndpointer that also accepts None (NULL)."""

    class _Opt(ptr_type):
        @classmethod
        def from_param(cls, obj):
            if obj is None:
                return None
            return ptr_type.from_param(obj)

    return _Opt


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = build()
    L = ctypes.CDLL(so)
    L.wd_oracle_num_threads.restype = _c_int
    L.wd_oracle_set_num_threads.argtypes = [_c_int]
    L.wd_oracle_tag_continuous_step.restype = None
    L.wd_oracle_tag_continuous_step.argtypes = [
        _c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _i32p, _f32p, _c_float, _c_float,
        _f32p, _f32p, _c_float, _c_int, _f32p, _c_int, _i32p, _c_int, _f32p, _i32p,
        _opt(_f32p), _opt(_i32p), _i32p, _f32p, _f32p, _i32p, _c_float, _c_float,
        _c_float, _c_float, _i32p, _i32p, _c_int, _c_int,
    ]
    L.wd_oracle_tag_gridworld_step.restype = None
    L.wd_oracle_tag_gridworld_step.argtypes = [
        _c_int, _c_int, _i32p, _i32p, _i32p, _i32p, _f32p, _f32p, _c_float, _c_float,
        _c_float, _c_float, _c_int, _c_int, _i32p, _c_int, _i32p,
    ]
    L.wd_oracle_mountain_car_step.restype = None
    L.wd_oracle_mountain_car_step.argtypes = (
        [_c_int, _f32p, _i32p, _i32p, _f32p, _f32p] + [_c_float] * 7 + [_i32p, _c_int])
    L.wd_oracle_continuous_mountain_car_step.restype = None
    L.wd_oracle_continuous_mountain_car_step.argtypes = (
        [_c_int, _f32p, _f32p, _i32p, _f32p, _f32p] + [_c_float] * 8 + [_i32p, _c_int])
    L.wd_oracle_pendulum_step.restype = None
    L.wd_oracle_pendulum_step.argtypes = [
        _c_int, _f32p, _f32p, _i32p, _f32p, _f32p, _i32p, _c_int]
    L.wd_oracle_acrobot_step.restype = None
    L.wd_oracle_acrobot_step.argtypes = [
        _c_int, _f32p, _i32p, _i32p, _f32p, _f32p, _i32p, _c_int]
    L.wd_oracle_cartpole_step.restype = None
    L.wd_oracle_cartpole_step.argtypes = [
        _c_int, _f32p, _i32p, _i32p, _f32p, _f32p, _c_float, _c_float, _c_float,
        _c_float, _c_float, _c_float, _c_float, _c_float, _c_float, _i32p, _c_int,
    ]
    L.wd_oracle_sample_actions.restype = None
    L.wd_oracle_sample_actions.argtypes = [
        _f32p, _i32p, _c_int, _opt(_f32p), _f32p, _c_long, _c_int, _c_int,
    ]
    L.wd_oracle_reset_when_done.restype = None
    L.wd_oracle_reset_when_done.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, _i32p, _c_int, _c_long, _c_int,
    ]
    L.wd_oracle_undo_done_and_reset_timestep.restype = None
    L.wd_oracle_undo_done_and_reset_timestep.argtypes = [_i32p, _i32p, _c_int, _c_int]
    L.wd_oracle_ou_process.restype = None
    L.wd_oracle_ou_process.argtypes = [
        _f32p, _f32p, _f32p, _f32p, _c_long, _c_float, _c_float, _c_float,
    ]
    L.wd_oracle_philox4x32_10.restype = None
    L.wd_oracle_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
    _lib = L
    return L


# --------------------------------------------------------------------------- #
# Thin, dict-of-arrays conveniences used by the parity tests.
# --------------------------------------------------------------------------- #
TC_STATE_KEYS = (
    "loc_x", "loc_y", "speed", "direction", "acceleration",
    "edge_hit_reward_penalty", "still_in_the_game", "num_runners",
    "nearest_neighbor_ids", "_done_", "_timestep_",
)


def tag_continuous_step(st, cfg, actions, obs=None, rewards=None, scratch=None):
    """Advance numpy state dict ``st`` in place by one step of the CUDA-path
    semantics.  ``cfg`` holds the per-env constants; returns (obs, rewards)."""
    E, N = st["loc_x"].shape
    K = int(cfg["num_other_agents_observed"])
    full = int(bool(cfg["use_full_observation"]))
    F = 7 * (N - 1) + 1 if full else 7 * K + 1
    if obs is None:
        obs = np.zeros((E, N, F), np.float32)
    if rewards is None:
        rewards = np.zeros((E, N), np.float32)
    nd = nid = None
    if scratch is not None:
        nd, nid = scratch
    lib().wd_oracle_tag_continuous_step(
        E, st["loc_x"], st["loc_y"], st["speed"], st["direction"], st["acceleration"],
        cfg["agent_types"], st["edge_hit_reward_penalty"],
        float(cfg["edge_hit_penalty"]), float(cfg["grid_length"]),
        cfg["acceleration_actions"], cfg["turn_actions"], float(cfg["max_speed"]), K,
        cfg["skill_levels"], int(bool(cfg["runner_exits_game_after_tagged"])),
        st["still_in_the_game"], full, obs,
        np.ascontiguousarray(actions, np.int32), nd, nid, st["nearest_neighbor_ids"],
        rewards, cfg["step_rewards"], st["num_runners"],
        float(cfg["distance_margin_for_reward"]), float(cfg["tag_reward_for_tagger"]),
        float(cfg["tag_penalty_for_runner"]),
        float(cfg["end_of_game_reward_for_runner"]), st["_done_"], st["_timestep_"],
        N, int(cfg["episode_length"]),
    )
    return obs, rewards


def philox4x32_10(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().wd_oracle_philox4x32_10(
        np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out
    )
    return out
