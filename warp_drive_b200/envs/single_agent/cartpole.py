"""CartPole-v1 as a single-agent WarpDrive env (10 000+ replicas, one thread each).

Host-side mirror of example_envs/single_agent/classic_control/cartpole/cartpole.py:14-126
and base.py:4-39 of the reference.  The reference delegates the CPU physics to
`gym.envs.classic_control.CartPoleEnv` (third-party, unpinned, not installed here -- see
SURVEY.md section 8c), so the CPU step below restates gym's Euler integrator in float64
with the constants the reference reads from the gym object (cartpole.py:69-81).  The
device step is wdb_cartpole_step, which follows the reference's numba kernel
(cartpole_step_numba.py:6-83).
"""
import math

import numpy as np

from warp_drive_b200.utils import spaces
from warp_drive_b200.utils.constants import Constants
from warp_drive_b200.utils.data_feed import DataFeed
from warp_drive_b200.utils.gpu_environment_context import CUDAEnvironmentContext

_OBSERVATIONS = Constants.OBSERVATIONS
_ACTIONS = Constants.ACTIONS
_REWARDS = Constants.REWARDS


class CartPolePhysics:
    """gym CartPoleEnv constants + Euler step (float64 state, float32 observations)."""

    gravity = 9.8
    masscart = 1.0
    masspole = 0.1
    total_mass = masspole + masscart
    length = 0.5  # half the pole length
    polemass_length = masspole * length
    force_mag = 10.0
    tau = 0.02
    theta_threshold_radians = 12 * 2 * math.pi / 360
    x_threshold = 2.4

    def __init__(self):
        self.state = None
        self._rng = np.random.default_rng()

    def reset(self, seed=None):
        if seed is not None:
            self._rng = np.random.default_rng(seed)
        self.state = self._rng.uniform(low=-0.05, high=0.05, size=(4,))
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if action == 1 else -self.force_mag
        costheta, sintheta = math.cos(theta), math.sin(theta)
        temp = (force + self.polemass_length * theta_dot ** 2 * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * costheta ** 2 / self.total_mass))
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = (x, x_dot, theta, theta_dot)
        terminated = bool(x < -self.x_threshold or x > self.x_threshold
                          or theta < -self.theta_threshold_radians
                          or theta > self.theta_threshold_radians)
        return np.array(self.state, dtype=np.float32), 1.0, terminated, False, {}


class SingleAgentEnv:
    def __init__(self, episode_length=500, env_backend="cpu", reset_pool_size=0, seed=None):
        self.num_agents = 1
        self.agents = {0: True}
        assert episode_length > 0
        self.episode_length = episode_length
        self.action_space = None
        self.observation_space = None
        self.timestep = None
        self.env_backend = env_backend
        # reset_pool_size < 2: every replica restarts from one fixed initial state
        self.reset_pool_size = reset_pool_size
        self.seed = seed


class ClassicControlCartPoleEnv(SingleAgentEnv):
    name = "ClassicControlCartPoleEnv"

    def __init__(self, episode_length, env_backend="cpu", reset_pool_size=0, seed=None):
        super().__init__(episode_length, env_backend, reset_pool_size, seed=seed)
        self.gym_env = CartPolePhysics()
        self.action_space = {0: spaces.Discrete(2)}
        high = np.array([4.8, np.finfo(np.float32).max, 0.42, np.finfo(np.float32).max],
                        dtype=np.float32)
        self.observation_space = {0: spaces.Box(-high, high, dtype=np.float32)}

    def step(self, action=None):
        self.timestep += 1
        assert isinstance(action, dict) and len(action) == 1
        state, reward, terminated, _, _ = self.gym_env.step(int(np.asarray(action[0]).reshape(-1)[0]))
        done = {"__all__": self.timestep >= self.episode_length or terminated}
        return {0: state}, {0: reward}, done, {}

    def reset(self):
        self.timestep = 0
        seed = self.seed if self.reset_pool_size < 2 else None
        state, _ = self.gym_env.reset(seed=seed)
        return {0: state}


_STEP_ARGS = [
    "state", _ACTIONS, "_done_", _REWARDS, _OBSERVATIONS, "gravity", "masspole",
    "total_mass", "length", "polemass_length", "force_mag", "tau",
    "theta_threshold_radians", "x_threshold", "_timestep_", ("episode_length", "meta"),
]


class CUDAClassicControlCartPoleEnv(ClassicControlCartPoleEnv, CUDAEnvironmentContext):
    def __init__(self, *args, **kwargs):
        ClassicControlCartPoleEnv.__init__(self, *args, **kwargs)
        CUDAEnvironmentContext.__init__(self)

    def get_data_dictionary(self):
        d = DataFeed()
        initial_state, _ = self.gym_env.reset(seed=self.seed)
        d.add_data(name="state", data=np.atleast_2d(initial_state),
                   save_copy_and_apply_at_reset=self.reset_pool_size < 2)
        g = self.gym_env
        d.add_data_list([
            ("gravity", g.gravity), ("masspole", g.masspole),
            ("total_mass", g.masspole + g.masscart), ("length", g.length),
            ("polemass_length", g.masspole * g.length), ("force_mag", g.force_mag),
            ("tau", g.tau), ("theta_threshold_radians", g.theta_threshold_radians),
            ("x_threshold", g.x_threshold),
        ])
        return d

    def get_reset_pool_dictionary(self):
        pool = DataFeed()
        if self.reset_pool_size >= 2:
            states = np.stack([np.atleast_2d(self.gym_env.reset(seed=None)[0])
                               for _ in range(self.reset_pool_size)], axis=0)
            assert states.ndim == 3 and states.shape[2] == 4
            pool.add_pool_for_reset(name="state_reset_pool", data=states,
                                    reset_target="state")
        return pool

    def step(self, actions=None):
        self.timestep += 1
        args = self.cuda_step_function_feed(_STEP_ARGS)
        if self.env_backend == "cpu":
            raise Exception("CUDAClassicControlCartPoleEnv expects a device env_backend")
        self.cuda_step[self.cuda_function_manager.grid,
                       self.cuda_function_manager.block](*args)
