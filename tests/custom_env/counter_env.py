"""CounterEnv: the host side of tests/custom_env/counter_env_step.cu, written like a
reference custom env (CPU step in NumPy + CUDAEnvironmentContext twin whose step() calls
`self.cuda_step(*self.cuda_step_function_feed(args), block=..., grid=...)` pycuda-style,
cf. example_envs/tag_gridworld/tag_gridworld.py:353-374)."""
import numpy as np

from warp_drive_b200.utils import spaces
from warp_drive_b200.utils.constants import Constants
from warp_drive_b200.utils.data_feed import DataFeed
from warp_drive_b200.utils.gpu_environment_context import CUDAEnvironmentContext

STEP_TABLE = np.array([0, 1, 2, 5], dtype=np.int32)


class CounterEnv:
    name = "CounterEnv"

    def __init__(self, num_agents=5, episode_length=20, limit=60, env_backend="cpu"):
        self.num_agents = num_agents
        self.episode_length = episode_length
        self.limit = limit
        self.env_backend = env_backend
        self.action_space = {a: spaces.Discrete(4) for a in range(num_agents)}
        self.observation_space = None
        self.timestep = 0
        self.counters = np.zeros(num_agents, np.int32)

    def _obs(self):
        total = float(self.counters.sum())
        return {a: np.array([self.counters[a], total, self.timestep / self.episode_length],
                            np.float32) for a in range(self.num_agents)}

    def reset(self):
        self.timestep = 0
        self.counters = np.zeros(self.num_agents, np.int32)
        return self._obs()

    def step(self, actions=None):
        self.timestep += 1
        for a in range(self.num_agents):
            self.counters[a] += STEP_TABLE[int(actions[a])]
        rew = {a: float(self.counters[a]) for a in range(self.num_agents)}
        done = {"__all__": self.timestep == self.episode_length
                or int(self.counters.sum()) >= self.limit}
        return self._obs(), rew, done, {}


class CUDACounterEnv(CounterEnv, CUDAEnvironmentContext):
    def __init__(self, *args, **kwargs):
        CounterEnv.__init__(self, *args, **kwargs)
        CUDAEnvironmentContext.__init__(self)

    def get_data_dictionary(self):
        d = DataFeed()
        d.add_data(name="counters", data=np.zeros(self.num_agents, np.int32),
                   save_copy_and_apply_at_reset=True)
        d.add_data(name="env_sum", data=0, save_copy_and_apply_at_reset=True)  # -> [n_envs]
        d.add_data(name="limit", data=self.limit)
        d.add_data(name="reset_value", data=7)     # argument of CudaCounterEnvReset
        return d

    def step(self, actions=None):
        self.timestep += 1
        args = ["counters", Constants.ACTIONS, "_done_", Constants.REWARDS,
                Constants.OBSERVATIONS, "env_sum", "limit", "_timestep_", ("episode_length", "meta")]
        self.cuda_step(*self.cuda_step_function_feed(args),
                       block=self.cuda_function_manager.block,
                       grid=self.cuda_function_manager.grid)
