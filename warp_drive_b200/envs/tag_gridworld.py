"""TagGridWorld: N taggers chase one runner on an integer grid.

Host-side mirror of the reference classes TagGridWorld / CUDATagGridWorld /
CUDATagGridWorldWithResetPool (example_envs/tag_gridworld/tag_gridworld.py:22-475): same
constructor, same data-dictionary names, same positional argument list for the device
step (:353-368).  Agents 0..N-2 are taggers (type 0), agent N-1 is the runner (type 1).
Integer state: the device kernel (wdb_tag_gridworld_step) is bit-exact against the
reference kernel's golden vectors.
"""
import numpy as np

from warp_drive_b200.utils import spaces
from warp_drive_b200.utils.constants import Constants
from warp_drive_b200.utils.data_feed import DataFeed
from warp_drive_b200.utils.gpu_environment_context import CUDAEnvironmentContext

_OBSERVATIONS = Constants.OBSERVATIONS
_ACTIONS = Constants.ACTIONS
_REWARDS = Constants.REWARDS

# no-op, right, left, up, down (tag_gridworld.py:104)
STEP_ACTIONS = np.array([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1]])


class TagGridWorld:
    name = "TagGridWorld"

    def __init__(self, num_taggers=10, grid_length=10, episode_length=100,
                 starting_location_x=None, starting_location_y=None, seed=None,
                 wall_hit_penalty=0.1, tag_reward_for_tagger=10.0,
                 tag_penalty_for_runner=2.0, step_cost_for_tagger=0.01,
                 use_full_observation=True, env_backend="cpu"):
        assert num_taggers > 0 and episode_length > 0
        self.num_taggers = num_taggers
        self.num_agents = num_taggers + 1
        self.episode_length = episode_length
        self.grid_length = grid_length
        self.np_random = np.random
        if seed is not None:
            self.seed(seed)
        N = self.num_agents
        self.agent_type = {a: int(a == N - 1) for a in range(N)}
        self.taggers = {a: True for a in range(N - 1)}
        self.runners = {N - 1: True}
        if starting_location_x is None:
            assert starting_location_y is None
            # taggers start in the centre, the runner in the corner (0, 0)
            starting_location_x = int(0.5 * grid_length) * np.ones(N)
            starting_location_y = int(0.5 * grid_length) * np.ones(N)
            starting_location_x[-1] = 0
            starting_location_y[-1] = 0
        assert len(starting_location_x) == N and len(starting_location_y) == N
        self.starting_location_x = np.asarray(starting_location_x)
        self.starting_location_y = np.asarray(starting_location_y)
        self.step_actions = STEP_ACTIONS.copy()
        self.observation_space = None  # set by EnvWrapper
        self.action_space = {a: spaces.Discrete(len(self.step_actions)) for a in range(N)}
        self.timestep = None
        self.global_state = None
        self.wall_hit_penalty = wall_hit_penalty
        self.tag_reward_for_tagger = tag_reward_for_tagger
        self.tag_penalty_for_runner = tag_penalty_for_runner
        self.step_cost_for_tagger = step_cost_for_tagger
        self.use_full_observation = use_full_observation
        self.env_backend = env_backend

    def seed(self, seed=None):
        self.np_random.seed(seed)
        return [seed]

    def reset(self):
        self.timestep = 0
        shape = (self.episode_length + 1, self.num_agents)
        self.global_state = {
            "loc_x": np.zeros(shape, dtype=np.int32),
            "loc_y": np.zeros(shape, dtype=np.int32),
        }
        self.global_state["loc_x"][0] = self.starting_location_x
        self.global_state["loc_y"][0] = self.starting_location_y
        return self.generate_observation()

    def generate_observation(self):
        t, N, B = self.timestep, self.num_agents, self.grid_length
        x = self.global_state["loc_x"][t].astype(np.float32) / np.float32(B)
        y = self.global_state["loc_y"][t].astype(np.float32) / np.float32(B)
        time = np.float32(t) / np.float32(self.episode_length)
        types = np.array([self.agent_type[a] for a in range(N)], dtype=np.float32)
        obs = {}
        if self.use_full_observation:
            for a in range(N):
                me = np.zeros(N, dtype=np.float32)
                me[a] = 1
                obs[a] = np.concatenate([x, y, types, me, [time]]).astype(np.float32)
            return obs
        xi, yi = self.global_state["loc_x"][t], self.global_state["loc_y"][t]
        d2 = (xi[:-1] - xi[-1]) ** 2 + (yi[:-1] - yi[-1]) ** 2
        closest = int(np.argmin(d2))
        for a in range(N):
            other = N - 1 if a < N - 1 else closest
            obs[a] = np.array([x[a], y[a], x[other], y[other], types[a], time],
                              dtype=np.float32)
        return obs

    def step(self, actions=None):
        self.timestep += 1
        assert isinstance(actions, dict) and len(actions) == self.num_agents
        t, N, B = self.timestep, self.num_agents, self.grid_length
        move = self.step_actions[[int(actions[a]) for a in range(N)]]
        x = self.global_state["loc_x"][t - 1] + move[:, 0]
        y = self.global_state["loc_y"][t - 1] + move[:, 1]
        cx, cy = np.clip(x, 0, B), np.clip(y, 0, B)
        rew = -self.wall_hit_penalty * ((x != cx).astype(np.float64) + (y != cy))
        self.global_state["loc_x"][t] = cx
        self.global_state["loc_y"][t] = cy
        tag = bool(((cx[:-1] == cx[-1]) & (cy[:-1] == cy[-1])).any())
        if tag:
            rew[:-1] += self.tag_reward_for_tagger
            rew[-1] -= self.tag_penalty_for_runner
        else:
            rew[:-1] -= self.step_cost_for_tagger
            rew[-1] += self.step_cost_for_tagger
        obs = self.generate_observation()
        done = {"__all__": t >= self.episode_length or tag}
        return obs, {a: rew[a] for a in range(N)}, done, {}


_STEP_ARGS = [
    "loc_x", "loc_y", _ACTIONS, "_done_", _REWARDS, _OBSERVATIONS, "wall_hit_penalty",
    "tag_reward_for_tagger", "tag_penalty_for_runner", "step_cost_for_tagger",
    "use_full_observation", "world_boundary", "_timestep_", ("episode_length", "meta"),
]


class CUDATagGridWorld(TagGridWorld, CUDAEnvironmentContext):
    """Device version: step() launches wdb_tag_gridworld_step.  The action table is a
    shared constant named kIndexToActionArr, initialised like in the reference
    (add_shared_constants + initialize_shared_constants); if the caller has not done it,
    it is done here on first use."""

    def __init__(self, *args, **kwargs):
        TagGridWorld.__init__(self, *args, **kwargs)
        CUDAEnvironmentContext.__init__(self)

    def get_data_dictionary(self):
        d = DataFeed()
        for key in ("loc_x", "loc_y"):
            d.add_data(name=key, data=self.global_state[key][0],
                       save_copy_and_apply_at_reset=True,
                       log_data_across_episode=True)
        d.add_data(name="wall_hit_penalty", data=self.wall_hit_penalty)
        d.add_data(name="tag_reward_for_tagger", data=self.tag_reward_for_tagger)
        d.add_data(name="tag_penalty_for_runner", data=self.tag_penalty_for_runner)
        d.add_data(name="step_cost_for_tagger", data=self.step_cost_for_tagger)
        d.add_data(name="use_full_observation", data=self.use_full_observation)
        d.add_data(name="world_boundary", data=self.grid_length)
        return d

    def _ensure_action_table(self):
        fm, dm = self.cuda_function_manager, self.cuda_data_manager
        if "kIndexToActionArr" not in fm._shared_tensors:
            if "kIndexToActionArr" not in dm._shared_constants:
                dm.add_shared_constants({"kIndexToActionArr": self.step_actions})
            fm.initialize_shared_constants(dm, constant_names=["kIndexToActionArr"])

    def step(self, actions=None):
        self.timestep += 1
        self._ensure_action_table()
        args = self.cuda_step_function_feed(_STEP_ARGS)
        if self.env_backend == "numba":
            self.cuda_step[self.cuda_function_manager.grid,
                           self.cuda_function_manager.block](*args)
        else:
            self.cuda_step(*args, block=self.cuda_function_manager.block,
                           grid=self.cuda_function_manager.grid)


class CUDATagGridWorldWithResetPool(CUDATagGridWorld):
    """Same env with random start positions drawn from a pool at every device reset
    (reference: tag_gridworld.py:383-475)."""

    def __init__(self, *args, reset_pool_size=100, **kwargs):
        super().__init__(*args, **kwargs)
        self.reset_pool_size = reset_pool_size

    def get_data_dictionary(self):
        d = DataFeed()
        for key in ("loc_x", "loc_y"):
            d.add_data(name=key, data=self.global_state[key][0],
                       save_copy_and_apply_at_reset=False)
        d.add_data(name="wall_hit_penalty", data=self.wall_hit_penalty)
        d.add_data(name="tag_reward_for_tagger", data=self.tag_reward_for_tagger)
        d.add_data(name="tag_penalty_for_runner", data=self.tag_penalty_for_runner)
        d.add_data(name="step_cost_for_tagger", data=self.step_cost_for_tagger)
        d.add_data(name="use_full_observation", data=self.use_full_observation)
        d.add_data(name="world_boundary", data=self.grid_length)
        return d

    def get_reset_pool_dictionary(self):
        N, B = self.num_agents, self.grid_length
        pool = DataFeed()
        pool.add_pool_for_reset(
            name="loc_x_reset_pool",
            data=self.np_random.randint(0, B + 1, (self.reset_pool_size, N)).astype(np.int32),
            reset_target="loc_x")
        pool.add_pool_for_reset(
            name="loc_y_reset_pool",
            data=self.np_random.randint(0, B + 1, (self.reset_pool_size, N)).astype(np.int32),
            reset_target="loc_y")
        return pool
