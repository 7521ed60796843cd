"""Rollout placeholders on the device: step-wise observations / sampled_actions / rewards
and the per-policy training batches the rollout kernels push into.

Names, shapes and dtypes follow the reference's create_and_push_data_placeholders
(warp_drive/training/utils/data_loader.py:30-240) because they ARE the interface between
env step kernels, the sampler, the models and the trainer:

  observations[_{policy}][_{key}]   f32 [E, N, *obs]   (save_copy_and_apply_at_reset)
  sampled_actions[_{policy}]        i32 [E, N, n_heads]  (f32 [E, N, dim] for Box actions)
  sampled_actions_{k}[_{policy}]    i32 [E, N, 1]       one per head of a MultiDiscrete
  rewards[_{policy}]                f32 [E, N]
  processed_observations_batch_{p}  f32 [T, E, Np, F]
  sampled_actions_batch_{p}         i32 [T, E, Np, n_heads]
  rewards_batch_{p}                 f32 [T, E, Np]
  done_flags_batch                  i32 [T, E]
"""
import numpy as np

from warp_drive_b200.utils.constants import Constants
from warp_drive_b200.utils.data_feed import DataFeed
from warp_drive_b200.utils.spaces import Box, Dict, Discrete, MultiDiscrete

_OBSERVATIONS = Constants.OBSERVATIONS
_ACTIONS = Constants.ACTIONS
_REWARDS = Constants.REWARDS
_DONE_FLAGS = Constants.DONE_FLAGS
_PROCESSED_OBSERVATIONS = Constants.PROCESSED_OBSERVATIONS


def get_flattened_obs_size(observation_space):
    """Number of float features a model sees for one agent (action masks excluded)."""
    if isinstance(observation_space, Box):
        return int(np.prod(observation_space.shape)) if observation_space.shape else 1
    if isinstance(observation_space, Dict):
        return sum(get_flattened_obs_size(v) for k, v in observation_space.items()
                   if k != Constants.ACTION_MASK)
    raise NotImplementedError("Observation space must be of Box or Dict type")


def action_head_sizes(action_space):
    """([head sizes], is_continuous) of one agent's action space."""
    if isinstance(action_space, Discrete):
        return [int(action_space.n)], False
    if isinstance(action_space, MultiDiscrete):
        return [int(n) for n in action_space.nvec], False
    if isinstance(action_space, Box):
        return [1] * int(action_space.shape[0]), True
    raise NotImplementedError("Action spaces can be of type Discrete, MultiDiscrete or Box")


def validate_policy_map(env_wrapper, policy_tag_to_agent_id_map):
    """Every agent belongs to exactly one policy; a single policy covers everyone by
    default (reference data_loader.py:243-271)."""
    n = env_wrapper.n_agents
    if policy_tag_to_agent_id_map is None:
        return {"shared": list(range(n))}
    seen = sorted(a for ids in policy_tag_to_agent_id_map.values() for a in ids)
    assert seen == list(range(n)), "every agent id must map to exactly one policy"
    return policy_tag_to_agent_id_map


def _check_same_spaces(env_wrapper, agent_ids):
    first = agent_ids[0]
    osp, asp = env_wrapper.env.observation_space, env_wrapper.env.action_space
    for a in agent_ids[1:]:
        assert get_flattened_obs_size(osp[a]) == get_flattened_obs_size(osp[first]), (
            "agents sharing a placeholder need identical observation shapes")
        assert action_head_sizes(asp[a]) == action_head_sizes(asp[first]), (
            "agents sharing a placeholder need identical action spaces")


def _push_step_placeholders(env_wrapper, action_sampler, agent_ids, suffix, obs_dim):
    dm = env_wrapper.cuda_data_manager
    E = env_wrapper.n_envs
    n = len(agent_ids)
    first = agent_ids[0]
    obs = env_wrapper.obs_at_reset()
    ospace = env_wrapper.env.observation_space[first]
    tensors = DataFeed()

    def stacked(getter):
        arr = np.stack([np.asarray(getter(obs[a])) for a in agent_ids], axis=0)
        if obs_dim == "last" and arr.ndim > 1:
            arr = np.moveaxis(arr, 0, -1)
        return np.broadcast_to(arr, (E,) + arr.shape).copy()

    if isinstance(ospace, Box):
        tensors.add_data(name=f"{_OBSERVATIONS}{suffix}", data=stacked(lambda o: o),
                         save_copy_and_apply_at_reset=True)
    elif isinstance(ospace, Dict):
        for key in ospace:
            tensors.add_data(name=f"{_OBSERVATIONS}{suffix}_{key}",
                             data=stacked(lambda o, k=key: o[k]),
                             save_copy_and_apply_at_reset=True)
    else:
        raise NotImplementedError("Observation space must be of Box or Dict type")

    heads, continuous = action_head_sizes(env_wrapper.env.action_space[first])
    act_dtype = np.float32 if continuous else np.int32
    tensors.add_data(name=f"{_ACTIONS}{suffix}", data=np.zeros((E, n, len(heads)), act_dtype))
    if len(heads) > 1:
        for k in range(len(heads)):
            tensors.add_data(name=f"{_ACTIONS}_{k}{suffix}", data=np.zeros((E, n, 1), act_dtype))
    tensors.add_data(name=f"{_REWARDS}{suffix}", data=np.zeros((E, n), np.float32))
    dm.push_data_to_device(tensors, torch_accessible=True)

    if action_sampler is not None:
        if len(heads) == 1:
            action_sampler.register_actions(dm, f"{_ACTIONS}{suffix}", heads[0],
                                            is_deterministic=continuous)
        else:
            for k, size in enumerate(heads):
                action_sampler.register_actions(dm, f"{_ACTIONS}_{k}{suffix}", size,
                                                is_deterministic=continuous)


def create_and_push_data_placeholders(
        env_wrapper=None, action_sampler=None, policy_tag_to_agent_id_map=None,
        create_separate_placeholders_for_each_policy=False,
        obs_dim_corresponding_to_num_agents="first", training_batch_size_per_env=None,
        push_data_batch_placeholders=True):
    assert env_wrapper is not None and env_wrapper.env_backend != "cpu"
    policy_map = validate_policy_map(env_wrapper, policy_tag_to_agent_id_map)
    if push_data_batch_placeholders:
        assert training_batch_size_per_env and training_batch_size_per_env > 0
    dm = env_wrapper.cuda_data_manager
    E = env_wrapper.n_envs

    if create_separate_placeholders_for_each_policy:
        assert len(policy_map) > 1
        for tag, ids in policy_map.items():
            _check_same_spaces(env_wrapper, ids)
            _push_step_placeholders(env_wrapper, action_sampler, ids, f"_{tag}",
                                    obs_dim_corresponding_to_num_agents)
    else:
        ids = list(range(env_wrapper.n_agents))
        _check_same_spaces(env_wrapper, ids)
        _push_step_placeholders(env_wrapper, action_sampler, ids, "",
                                obs_dim_corresponding_to_num_agents)

    T = training_batch_size_per_env
    batches = DataFeed()
    if T is not None and T >= 1:
        for tag, ids in policy_map.items():
            F = get_flattened_obs_size(env_wrapper.env.observation_space[ids[0]])
            batches.add_data(name=f"{_PROCESSED_OBSERVATIONS}_batch_{tag}",
                             data=np.zeros((T, E, len(ids), F), np.float32))
    if push_data_batch_placeholders:
        for tag, ids in policy_map.items():
            heads, continuous = action_head_sizes(env_wrapper.env.action_space[ids[0]])
            batches.add_data(name=f"{_ACTIONS}_batch_{tag}",
                             data=np.zeros((T, E, len(ids), len(heads)),
                                           np.float32 if continuous else np.int32))
            batches.add_data(name=f"{_REWARDS}_batch_{tag}",
                             data=np.zeros((T, E, len(ids)), np.float32))
        batches.add_data(name=f"{_DONE_FLAGS}_batch", data=np.zeros((T, E), np.int32))
    dm.push_data_to_device(batches, torch_accessible=True)
    return policy_map
