"""MountainCar, ContinuousMountainCar, Acrobot and Pendulum as single-agent WarpDrive envs.

Host-side mirror of example_envs/single_agent/classic_control/{mountain_car,
continuous_mountain_car,acrobot,pendulum}/*.py of the reference (class names, `name`
attributes, constructor arguments, data dictionaries, reset pools and the positional `args`
lists of step() are the reference's).  The reference delegates the CPU physics to
`gym.envs.classic_control.*` (third-party, unpinned, not installed here -- SURVEY.md
section 8c); the `*Physics` classes below restate gym's published integrators in float64
with gym's constants, exactly as cartpole.py does for CartPole.  The device steps are the
`wdb_*_step` kernels of csrc/wdb_classic_control.cu, which follow the reference's numba
kernels (the `*_step_numba.py` files), not gym.
"""
import math

import numpy as np

from warp_drive_b200.envs.single_agent.cartpole import SingleAgentEnv
from warp_drive_b200.utils import spaces
from warp_drive_b200.utils.constants import Constants
from warp_drive_b200.utils.data_feed import DataFeed
from warp_drive_b200.utils.gpu_environment_context import CUDAEnvironmentContext

_OBSERVATIONS = Constants.OBSERVATIONS
_ACTIONS = Constants.ACTIONS
_REWARDS = Constants.REWARDS


class _Physics:
    def __init__(self):
        self.state = None
        self._rng = np.random.default_rng()

    def _seed(self, seed):
        if seed is not None:
            self._rng = np.random.default_rng(seed)


class MountainCarPhysics(_Physics):
    """gym MountainCarEnv (MountainCar-v0): constants and Euler step."""

    min_position, max_position, max_speed = -1.2, 0.6, 0.07
    goal_position, goal_velocity = 0.5, 0.0
    force, gravity = 0.001, 0.0025

    @property
    def action_space(self):
        return spaces.Discrete(3)

    @property
    def observation_space(self):
        low = np.array([self.min_position, -self.max_speed], dtype=np.float32)
        high = np.array([self.max_position, self.max_speed], dtype=np.float32)
        return spaces.Box(low, high, dtype=np.float32)

    def reset(self, seed=None):
        self._seed(seed)
        self.state = np.array([self._rng.uniform(low=-0.6, high=-0.4), 0.0])
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        position, velocity = self.state
        velocity += (int(action) - 1) * self.force + math.cos(3 * position) * (-self.gravity)
        velocity = float(np.clip(velocity, -self.max_speed, self.max_speed))
        position += velocity
        position = float(np.clip(position, self.min_position, self.max_position))
        if position == self.min_position and velocity < 0:
            velocity = 0.0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        self.state = (position, velocity)
        return np.array(self.state, dtype=np.float32), -1.0, terminated, False, {}


class ContinuousMountainCarPhysics(_Physics):
    """gym Continuous_MountainCarEnv (MountainCarContinuous-v0)."""

    min_action, max_action = -1.0, 1.0
    min_position, max_position, max_speed = -1.2, 0.6, 0.07
    goal_position, goal_velocity = 0.45, 0.0
    power = 0.0015

    @property
    def action_space(self):
        return spaces.Box(self.min_action, self.max_action, shape=(1,), dtype=np.float32)

    @property
    def observation_space(self):
        low = np.array([self.min_position, -self.max_speed], dtype=np.float32)
        high = np.array([self.max_position, self.max_speed], dtype=np.float32)
        return spaces.Box(low, high, dtype=np.float32)

    def reset(self, seed=None):
        self._seed(seed)
        self.state = np.array([self._rng.uniform(low=-0.6, high=-0.4), 0.0])
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        a = float(np.asarray(action).reshape(-1)[0])
        position, velocity = float(self.state[0]), float(self.state[1])
        force = min(max(a, self.min_action), self.max_action)
        velocity += force * self.power - 0.0025 * math.cos(3 * position)
        velocity = min(max(velocity, -self.max_speed), self.max_speed)
        position += velocity
        position = min(max(position, self.min_position), self.max_position)
        if position == self.min_position and velocity < 0:
            velocity = 0.0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        reward = 100.0 if terminated else 0.0
        reward -= math.pow(a, 2) * 0.1
        self.state = np.array([position, velocity], dtype=np.float32)
        return self.state, reward, terminated, False, {}


class PendulumPhysics(_Physics):
    """gym PendulumEnv(g=9.81) (Pendulum-v1)."""

    max_speed, max_torque, dt, m, l = 8, 2.0, 0.05, 1.0, 1.0

    def __init__(self, g=9.81):
        super().__init__()
        self.g = g

    @property
    def action_space(self):
        return spaces.Box(-self.max_torque, self.max_torque, shape=(1,), dtype=np.float32)

    @property
    def observation_space(self):
        high = np.array([1.0, 1.0, self.max_speed], dtype=np.float32)
        return spaces.Box(-high, high, dtype=np.float32)

    def _get_obs(self):
        theta, thetadot = self.state
        return np.array([np.cos(theta), np.sin(theta), thetadot], dtype=np.float32)

    def reset(self, seed=None):
        self._seed(seed)
        high = np.array([np.pi, 1.0])
        self.state = self._rng.uniform(low=-high, high=high)
        return self._get_obs(), {}

    def step(self, u):
        th, thdot = self.state
        u = float(np.clip(np.asarray(u, dtype=np.float64).reshape(-1)[0],
                          -self.max_torque, self.max_torque))
        angle = ((th + np.pi) % (2 * np.pi)) - np.pi
        costs = angle ** 2 + 0.1 * thdot ** 2 + 0.001 * (u ** 2)
        newthdot = thdot + (3 * self.g / (2 * self.l) * np.sin(th)
                            + 3.0 / (self.m * self.l ** 2) * u) * self.dt
        newthdot = np.clip(newthdot, -self.max_speed, self.max_speed)
        newth = th + newthdot * self.dt
        self.state = np.array([newth, newthdot])
        return self._get_obs(), -costs, False, False, {}


class AcrobotPhysics(_Physics):
    """gym AcrobotEnv (Acrobot-v1): "book" dynamics, RK4 over dt = 0.2, no torque noise."""

    dt = 0.2
    LINK_LENGTH_1 = LINK_LENGTH_2 = 1.0
    LINK_MASS_1 = LINK_MASS_2 = 1.0
    LINK_COM_POS_1 = LINK_COM_POS_2 = 0.5
    LINK_MOI = 1.0
    MAX_VEL_1, MAX_VEL_2 = 4 * math.pi, 9 * math.pi
    AVAIL_TORQUE = (-1.0, 0.0, +1.0)

    @property
    def action_space(self):
        return spaces.Discrete(3)

    @property
    def observation_space(self):
        high = np.array([1.0, 1.0, 1.0, 1.0, self.MAX_VEL_1, self.MAX_VEL_2], dtype=np.float32)
        return spaces.Box(-high, high, dtype=np.float32)

    def _get_ob(self):
        s = self.state
        return np.array([math.cos(s[0]), math.sin(s[0]), math.cos(s[1]), math.sin(s[1]),
                         s[2], s[3]], dtype=np.float32)

    def reset(self, seed=None):
        self._seed(seed)
        self.state = self._rng.uniform(low=-0.1, high=0.1, size=(4,)).astype(np.float32)
        return self._get_ob(), {}

    def _dsdt(self, s, a):
        m1, m2, l1 = self.LINK_MASS_1, self.LINK_MASS_2, self.LINK_LENGTH_1
        lc1, lc2 = self.LINK_COM_POS_1, self.LINK_COM_POS_2
        i1 = i2 = self.LINK_MOI
        g = 9.8
        theta1, theta2, dtheta1, dtheta2 = s
        d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * math.cos(theta2)) + i1 + i2
        d2 = m2 * (lc2 ** 2 + l1 * lc2 * math.cos(theta2)) + i2
        phi2 = m2 * lc2 * g * math.cos(theta1 + theta2 - math.pi / 2.0)
        phi1 = (-m2 * l1 * lc2 * dtheta2 ** 2 * math.sin(theta2)
                - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * math.sin(theta2)
                + (m1 * lc1 + m2 * l1) * g * math.cos(theta1 - math.pi / 2) + phi2)
        ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * dtheta1 ** 2 * math.sin(theta2) - phi2) / (
            m2 * lc2 ** 2 + i2 - d2 ** 2 / d1)
        ddtheta1 = -(d2 * ddtheta2 + phi1) / d1
        return np.array([dtheta1, dtheta2, ddtheta1, ddtheta2], dtype=np.float64)

    @staticmethod
    def _wrap(x, m, M):
        diff = M - m
        while x > M:
            x = x - diff
        while x < m:
            x = x + diff
        return x

    def step(self, a):
        s = np.asarray(self.state, dtype=np.float64)
        torque = self.AVAIL_TORQUE[int(a)]
        dt, dt2 = self.dt, self.dt / 2.0
        k1 = self._dsdt(s, torque)
        k2 = self._dsdt(s + dt2 * k1, torque)
        k3 = self._dsdt(s + dt2 * k2, torque)
        k4 = self._dsdt(s + dt * k3, torque)
        ns = s + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
        ns[0] = self._wrap(ns[0], -math.pi, math.pi)
        ns[1] = self._wrap(ns[1], -math.pi, math.pi)
        ns[2] = min(max(ns[2], -self.MAX_VEL_1), self.MAX_VEL_1)
        ns[3] = min(max(ns[3], -self.MAX_VEL_2), self.MAX_VEL_2)
        self.state = ns
        terminated = bool(-math.cos(ns[0]) - math.cos(ns[1] + ns[0]) > 1.0)
        reward = -1.0 if not terminated else 0.0
        return self._get_ob(), reward, terminated, False, {}


# --------------------------------------------------------------------------- env classes
class _ClassicControlEnv(SingleAgentEnv):
    """CPU env: same step()/reset() contract as the reference's ClassicControl*Env."""

    _physics = None

    def __init__(self, episode_length, env_backend="cpu", reset_pool_size=0, seed=None):
        super().__init__(episode_length, env_backend, reset_pool_size, seed=seed)
        self.gym_env = self._physics()
        self.action_space = {0: self.gym_env.action_space}
        self.observation_space = {0: self.gym_env.observation_space}

    def step(self, action=None):
        self.timestep += 1
        assert isinstance(action, dict) and len(action) == 1
        obs, reward, terminated, _, _ = self.gym_env.step(action[0])
        done = {"__all__": self.timestep >= self.episode_length or terminated}
        return {0: obs}, {0: reward}, done, {}

    def reset(self):
        self.timestep = 0
        seed = self.seed if self.reset_pool_size < 2 else None
        obs, _ = self.gym_env.reset(seed=seed)
        return {0: obs}


class _CUDAClassicControlEnv(CUDAEnvironmentContext):
    """Device env: `state` (+ an optional reset pool) lives in the data manager, step() is one
    kernel launch with the reference's positional argument list."""

    _constants = ()        # names of the gym attributes pushed as float32 scalars
    _state_dim = 2
    _state_from_reset = True   # MountainCar*: reset() returns the state; Acrobot/Pendulum:
                               # reset() returns the observation, the state is gym_env.state

    def _initial_state(self, seed):
        first, _ = self.gym_env.reset(seed=seed)
        return first if self._state_from_reset else self.gym_env.state

    def get_data_dictionary(self):
        d = DataFeed()
        d.add_data(name="state", data=np.atleast_2d(self._initial_state(self.seed)),
                   save_copy_and_apply_at_reset=self.reset_pool_size < 2)
        if self._constants:
            d.add_data_list([(k, getattr(self.gym_env, k)) for k in self._constants])
        return d

    def get_tensor_dictionary(self):
        return DataFeed()

    def get_reset_pool_dictionary(self):
        pool = DataFeed()
        if self.reset_pool_size >= 2:
            states = np.stack([np.atleast_2d(self._initial_state(None))
                               for _ in range(self.reset_pool_size)], axis=0)
            assert states.ndim == 3 and states.shape[2] == self._state_dim
            pool.add_pool_for_reset(name="state_reset_pool", data=states,
                                    reset_target="state")
        return pool

    def step(self, actions=None):
        self.timestep += 1
        args = (["state", _ACTIONS, "_done_", _REWARDS, _OBSERVATIONS] + list(self._constants)
                + ["_timestep_", ("episode_length", "meta")])
        if self.env_backend == "cpu":
            raise Exception(f"{type(self).__name__} expects a device env_backend")
        self.cuda_step[self.cuda_function_manager.grid,
                       self.cuda_function_manager.block](*self.cuda_step_function_feed(args))


def _device_env(cpu_cls, constants, state_dim, state_from_reset):
    class _Env(cpu_cls, _CUDAClassicControlEnv):
        _constants = tuple(constants)
        _state_dim = state_dim
        _state_from_reset = state_from_reset

        def __init__(self, *args, **kwargs):
            cpu_cls.__init__(self, *args, **kwargs)
            CUDAEnvironmentContext.__init__(self)

        step = _CUDAClassicControlEnv.step

    _Env.__name__ = _Env.__qualname__ = "CUDA" + cpu_cls.__name__
    return _Env


class ClassicControlMountainCarEnv(_ClassicControlEnv):
    name = "ClassicControlMountainCarEnv"
    _physics = MountainCarPhysics


class ClassicControlContinuousMountainCarEnv(_ClassicControlEnv):
    name = "ClassicControlContinuousMountainCarEnv"
    _physics = ContinuousMountainCarPhysics


class ClassicControlAcrobotEnv(_ClassicControlEnv):
    name = "ClassicControlAcrobotEnv"
    _physics = AcrobotPhysics


class ClassicControlPendulumEnv(_ClassicControlEnv):
    name = "ClassicControlPendulumEnv"
    _physics = PendulumPhysics


# argument lists: mountain_car.py:104-119, continuous_mountain_car.py:105-121,
# acrobot.py:91-99, pendulum.py:92-100 of the reference
CUDAClassicControlMountainCarEnv = _device_env(
    ClassicControlMountainCarEnv,
    ("min_position", "max_position", "max_speed", "goal_position", "goal_velocity", "force",
     "gravity"), 2, True)
CUDAClassicControlContinuousMountainCarEnv = _device_env(
    ClassicControlContinuousMountainCarEnv,
    ("min_action", "max_action", "min_position", "max_position", "max_speed",
     "goal_position", "goal_velocity", "power"), 2, True)
CUDAClassicControlAcrobotEnv = _device_env(ClassicControlAcrobotEnv, (), 4, False)
CUDAClassicControlPendulumEnv = _device_env(ClassicControlPendulumEnv, (), 2, False)
