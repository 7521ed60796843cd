"""Shared helpers for the parity tests (state construction from golden fixtures)."""
import numpy as np


def tc_cfg_from_fixture(fx):
    """Per-env constants of a tag_continuous fixture, typed like the device arrays."""
    cfg = {
        "agent_types": fx["init__agent_types"].astype(np.int32),
        "acceleration_actions": fx["init__acceleration_actions"].astype(np.float32),
        "turn_actions": fx["init__turn_actions"].astype(np.float32),
        "skill_levels": fx["init__skill_levels"].astype(np.float32),
        "step_rewards": fx["init__step_rewards"].astype(np.float32),
        "episode_length": int(fx["episode_length"]),
    }
    for key in ("grid_length", "edge_hit_penalty", "max_speed",
                "distance_margin_for_reward", "tag_reward_for_tagger",
                "tag_penalty_for_runner", "end_of_game_reward_for_runner"):
        cfg[key] = np.float32(fx[f"init__{key}"])
    for key in ("num_other_agents_observed", "use_full_observation",
                "runner_exits_game_after_tagged"):
        cfg[key] = int(fx[f"init__{key}"])
    return cfg


def tc_state_from_fixture(fx, n_envs):
    """Initial state replicated over n_envs (what EnvWrapper.reset_all_envs pushes)."""
    N = fx["init__loc_x"].shape[0]
    K = int(fx["init__num_other_agents_observed"])

    def rep(a, dtype):
        a = np.asarray(a).astype(dtype)
        return np.ascontiguousarray(np.broadcast_to(a, (n_envs,) + a.shape)).copy()

    st = {
        "loc_x": rep(fx["init__loc_x"], np.float32),
        "loc_y": rep(fx["init__loc_y"], np.float32),
        "speed": rep(fx["init__speed"], np.float32),
        "direction": rep(fx["init__direction"], np.float32),
        "acceleration": rep(fx["init__acceleration"], np.float32),
        "edge_hit_reward_penalty": rep(fx["init__edge_hit_reward_penalty"], np.float32),
        "still_in_the_game": rep(fx["init__still_in_the_game"], np.int32),
        "num_runners": np.full(n_envs, int(fx["init__num_runners"]), np.int32),
        "nearest_neighbor_ids": np.zeros((n_envs, N, K), np.int32),
        "_done_": np.zeros(n_envs, np.int32),
        "_timestep_": np.zeros(n_envs, np.int32),
    }
    return st


def copy_state(st):
    return {k: v.copy() for k, v in st.items()}


def close(a, b, tol=1e-5):
    """abs-or-rel closeness (the reference's own comparison rule, tightened from 1 %)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
