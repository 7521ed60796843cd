"""The whole T-step rollout of a discrete-action single-agent env in ONE launch
(`wdb_single_agent_rollout`, csrc/wdb_sa_rollout.cu): forward -> sample -> step ->
bookkeeping -> done-masked reset -> push to batch, one thread per env replica.

B200 re-design of TrainerBase._generate_rollout_batch (warp_drive/training/trainers/
trainer_base.py:383-428) for ClassicControl{CartPole,MountainCar,Acrobot}Env with a small
FullyConnected policy (BASELINE config 3): the reference issues ~50 launches and >= 5 host
synchronisations per timestep, this path one launch per T timesteps and no synchronisation.
"""
import ctypes

import torch

from warp_drive_b200 import lib as _lib
from warp_drive_b200.utils.constants import Constants

_OBSERVATIONS = Constants.OBSERVATIONS
_ACTIONS = Constants.ACTIONS
_REWARDS = Constants.REWARDS
_DONE_FLAGS = Constants.DONE_FLAGS
_PROCESSED_OBSERVATIONS = Constants.PROCESSED_OBSERVATIONS

# env class name -> (kind, scalar step arguments in kernel order, state_dim)
_ENVS = {
    "ClassicControlCartPoleEnv": (0, ("gravity", "masspole", "total_mass", "length",
                                      "polemass_length", "force_mag", "tau",
                                      "theta_threshold_radians", "x_threshold"), 4),
    "ClassicControlMountainCarEnv": (1, ("min_position", "max_position", "max_speed",
                                         "goal_position", "goal_velocity", "force", "gravity"), 2),
    "ClassicControlAcrobotEnv": (2, (), 4),
}


def _p(t):
    return _lib.ptr(t)


def _layers(model):
    """[(weight, bias)] of the hidden layers followed by the (single) action head."""
    out = [(model.fc[str(i)][0].weight, model.fc[str(i)][0].bias) for i in range(len(model.fc))]
    head = model.policy_head[0]
    out.append((head.weight, head.bias))
    return out


class FusedSingleAgentRollout:
    @staticmethod
    def eligible(env_wrapper, models, policy_map):
        env = env_wrapper.env
        if getattr(env, "name", "") not in _ENVS or env_wrapper.n_agents != 1:
            return False
        if len(policy_map) != 1 or len(models) != 1:
            return False
        model = next(iter(models.values()))
        try:
            if getattr(model, "is_deterministic", True) or len(model.output_dims) != 1:
                return False
            if getattr(model, "action_mask", None) is not None:
                return False
            if not 0 <= len(model.fc) <= 3:
                return False
            dims = [int(model.flattened_obs_size)] + [int(d) for d in model.fc_dims] + [
                int(model.output_dims[0])]
            arr = (ctypes.c_int * 5)(*(dims + [0] * (5 - len(dims))))
            return bool(_lib.load().wdb_single_agent_rollout_supported(len(model.fc), arr))
        except Exception:  # noqa: BLE001
            return False

    def __init__(self, env_wrapper, model, policy, sampler):
        self.dm = dm = env_wrapper.cuda_data_manager
        self.env_wrapper = env_wrapper
        self.model = model
        self.policy = policy
        self.sampler = sampler
        self.lib = _lib.load()
        self.E = env_wrapper.n_envs
        kind, names, state_dim = _ENVS[env_wrapper.env.name]
        r = self.r = _lib.SaRollout()
        r.env_kind, r.n_envs, r.state_dim = kind, self.E, state_dim
        r.episode_length = int(dm.meta_info("episode_length"))
        for i, name in enumerate(names):
            r.env_params[i] = float(dm.device_data(name))
        r.state = _p(dm.device_data("state"))
        r.observations = _p(dm.device_data(_OBSERVATIONS))
        r.done = _p(dm.device_data("_done_"))
        r.env_timestep = _p(dm.device_data("_timestep_"))
        r.rewards = _p(dm.device_data(_REWARDS))
        r.sampled_actions = _p(dm.device_data(_ACTIONS))
        layers = _layers(model)
        r.n_hidden = len(layers) - 1
        dims = [int(model.flattened_obs_size)] + [int(w.shape[0]) for w, _ in layers]
        for i, d in enumerate(dims):
            r.dims[i] = d
        self.F, self.A = dims[0], dims[-1]
        self._params = layers            # live nn.Parameters: optimizer steps update in place
        for i, (w, b) in enumerate(layers):
            assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
            r.w[i], r.b[i] = _p(w.data), _p(b.data)
        r.rng_state = _p(sampler.rng_state)
        r.reset_done_envs = 1
        self._table = None

    def set_bookkeeping(self, reward_running_sum, episodic_reward_sum, step_running_sum,
                        episodic_step_sum, num_completed_episodes):
        r = self.r
        r.reward_running_sum = _p(reward_running_sum[self.policy])
        r.episodic_reward_sum = _p(episodic_reward_sum[self.policy])
        r.step_running_sum = _p(step_running_sum)
        r.episodic_step_sum = _p(episodic_step_sum)
        r.num_completed = _p(num_completed_episodes)
        self._keep = (reward_running_sum, episodic_reward_sum, step_running_sum,
                      episodic_step_sum, num_completed_episodes)

    def _reset_table(self):
        resetter = self.env_wrapper.env_resetter
        table, n = resetter.build_table(self.dm)
        if self.dm.reset_target_to_pool:
            assert resetter._random_initialized, (
                "reset pools need init_reset_pool(data_manager, seed) first")
        return table, n, resetter._pool_rng

    def launch(self, n_steps, t0=0, record=True, uniforms=None, probs_out=None,
               use_argmax=False, reset_done_envs=True):
        """Run `n_steps` timesteps.  record=True pushes them into batch slots t0 .. t0+n_steps-1
        of the policy's training batch."""
        dm, r, p = self.dm, self.r, self.policy
        # parameters may have been re-allocated (load_state_dict keeps storage; .to() does not)
        for i, (w, b) in enumerate(self._params):
            r.w[i], r.b[i] = _p(w.data), _p(b.data)
        r.n_steps = int(n_steps)
        r.use_argmax = int(bool(use_argmax))
        r.uniforms = _p(uniforms)
        r.probs_batch = _p(probs_out)
        if record:
            obs_b = dm.data_on_device_via_torch(f"{_PROCESSED_OBSERVATIONS}_batch_{p}")
            assert t0 + n_steps <= obs_b.shape[0]
            r.obs_batch = _p(obs_b[t0])
            r.actions_batch = _p(dm.data_on_device_via_torch(f"{_ACTIONS}_batch_{p}")[t0])
            r.rewards_batch = _p(dm.data_on_device_via_torch(f"{_REWARDS}_batch_{p}")[t0])
            r.done_batch = _p(dm.data_on_device_via_torch(f"{_DONE_FLAGS}_batch")[t0])
        else:
            r.obs_batch = r.actions_batch = r.rewards_batch = r.done_batch = None
        table, n, pool_rng = self._reset_table()
        self._table = table
        r.reset_table = _p(table) if n > 0 else None
        r.n_reset_arrays = n
        r.pool_rng = _p(pool_rng)
        r.reset_done_envs = int(bool(reset_done_envs))
        _lib.check(self.lib.wdb_single_agent_rollout(_lib.stream_ptr(), ctypes.byref(r)),
                   "single_agent_rollout")
