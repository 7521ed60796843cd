"""TagContinuous: taggers chase runners on a continuous square grid.

Host-side mirror of the reference env class (example_envs/tag_continuous/
tag_continuous.py:28-887): same constructor arguments, same seeded initial state (the
np.random call order of :152-199 is reproduced so a given `seed` yields the reference's
taggers / positions / headings), same `get_data_dictionary` names, same positional
argument list for the device step (:806-840).  The device step is
wdb_tag_continuous_step (warp_drive_b200/csrc/wdb_tag_continuous.cu); the CPU step is a
vectorised NumPy implementation of the same rules, used by `env_backend="cpu"` and by
the CPU-vs-GPU consistency checker.
"""
import numpy as np

from warp_drive_b200.utils import spaces
from warp_drive_b200.utils.constants import Constants
from warp_drive_b200.utils.data_feed import DataFeed
from warp_drive_b200.utils.gpu_environment_context import CUDAEnvironmentContext

_OBSERVATIONS = Constants.OBSERVATIONS
_ACTIONS = Constants.ACTIONS
_REWARDS = Constants.REWARDS

_STATE_KEYS = ("loc_x", "loc_y", "speed", "direction", "acceleration")

# positional arguments of the device step == CudaTagContinuousStep's signature
_STEP_ARGS = [
    "loc_x", "loc_y", "speed", "direction", "acceleration", "agent_types",
    "edge_hit_reward_penalty", "edge_hit_penalty", "grid_length", "acceleration_actions",
    "turn_actions", "max_speed", "num_other_agents_observed", "skill_levels",
    "runner_exits_game_after_tagged", "still_in_the_game", "use_full_observation",
    _OBSERVATIONS, _ACTIONS, "neighbor_distances", "neighbor_ids_sorted_by_distance",
    "nearest_neighbor_ids", _REWARDS, "step_rewards", "num_runners",
    "distance_margin_for_reward", "tag_reward_for_tagger", "tag_penalty_for_runner",
    "end_of_game_reward_for_runner", "_done_", "_timestep_", ("n_agents", "meta"),
    ("episode_length", "meta"),
]


# arguments of the device step that carry one row per env replica
_PER_ENV_ARGS = {"loc_x", "loc_y", "speed", "direction", "acceleration",
                 "edge_hit_reward_penalty", "still_in_the_game", _OBSERVATIONS, _ACTIONS,
                 "neighbor_distances", "neighbor_ids_sorted_by_distance", "nearest_neighbor_ids",
                 _REWARDS, "num_runners", "_done_", "_timestep_"}


class TagContinuous(CUDAEnvironmentContext):
    name = "TagContinuous"

    def __init__(self, num_taggers=1, num_runners=10, grid_length=10.0, episode_length=100,
                 starting_location_x=None, starting_location_y=None,
                 starting_directions=None, seed=None, max_speed=1.0,
                 skill_level_runner=1.0, skill_level_tagger=1.0, max_acceleration=1.0,
                 min_acceleration=-1.0, max_turn=np.pi / 2, min_turn=-np.pi / 2,
                 num_acceleration_levels=10, num_turn_levels=10, edge_hit_penalty=-0.0,
                 use_full_observation=True, num_other_agents_observed=2,
                 tagging_distance=0.01, tag_reward_for_tagger=1.0,
                 step_penalty_for_tagger=-0.0, tag_penalty_for_runner=-1.0,
                 step_reward_for_runner=0.0, end_of_game_reward_for_runner=1.0,
                 runner_exits_game_after_tagged=True, env_backend="cpu",
                 allocate_reference_scratch=False):
        super().__init__()
        f32 = np.float32
        self.float_dtype, self.int_dtype = np.float32, np.int32
        self.eps = f32(1e-10)
        assert num_taggers > 0 and num_runners > 0 and episode_length > 0 and grid_length > 0
        self.num_taggers, self.num_runners = int(num_taggers), int(num_runners)
        self.num_agents = self.num_taggers + self.num_runners
        self.episode_length = int(episode_length)
        self.grid_length = f32(grid_length)
        self.grid_diagonal = self.grid_length * np.sqrt(2)
        assert edge_hit_penalty <= 0
        self.edge_hit_penalty = f32(edge_hit_penalty)

        # --- seeded initial conditions; the draw order below is the reference's
        self.np_random = np.random
        if seed is not None:
            self.seed(seed)
        N = self.num_agents
        tagger_ids = self.np_random.choice(np.arange(N), self.num_taggers, replace=False)
        is_tagger = np.zeros(N, dtype=bool)
        is_tagger[tagger_ids] = True
        self.agent_type = {a: int(is_tagger[a]) for a in range(N)}   # 1 tagger, 0 runner
        self.taggers = {a: True for a in range(N) if is_tagger[a]}
        self.runners = {a: True for a in range(N) if not is_tagger[a]}
        if starting_location_x is None:
            assert starting_location_y is None
            starting_location_x = self.grid_length * self.np_random.rand(N)
            starting_location_y = self.grid_length * self.np_random.rand(N)
        assert len(starting_location_x) == N and len(starting_location_y) == N
        self.starting_location_x = starting_location_x
        self.starting_location_y = starting_location_y
        if starting_directions is None:
            starting_directions = self.np_random.choice(
                [0, np.pi / 2, np.pi, np.pi * 3 / 2], N, replace=True)
        assert len(starting_directions) == N
        self.starting_directions = starting_directions
        self.starting_speeds = np.zeros(N, dtype=f32)
        self.starting_accelerations = np.zeros(N, dtype=f32)

        # --- action tables: index 0 is the no-op, then `levels` evenly spaced values
        assert num_acceleration_levels >= 0 and num_turn_levels >= 0
        self.max_speed = f32(max_speed)
        self.num_acceleration_levels = num_acceleration_levels
        self.num_turn_levels = num_turn_levels
        self.max_acceleration, self.min_acceleration = f32(max_acceleration), f32(min_acceleration)
        self.max_turn, self.min_turn = f32(max_turn), f32(min_turn)
        self.acceleration_actions = np.concatenate(
            [[0.0], np.linspace(self.min_acceleration, self.max_acceleration,
                                num_acceleration_levels)]).astype(f32)
        self.turn_actions = np.concatenate(
            [[0.0], np.linspace(self.min_turn, self.max_turn, num_turn_levels)]).astype(f32)

        types = is_tagger.astype(np.int64)
        self.skill_levels = [
            f32(skill_level_tagger) if is_tagger[a] else f32(skill_level_runner)
            for a in range(N)]
        self.runner_exits_game_after_tagged = runner_exits_game_after_tagged
        self.timestep = None
        self.global_state = None
        self.observation_space = None  # set by EnvWrapper
        self.action_space = {
            a: spaces.MultiDiscrete((len(self.acceleration_actions), len(self.turn_actions)))
            for a in range(N)}
        self.use_full_observation = use_full_observation
        assert num_other_agents_observed <= N
        self.num_other_agents_observed = num_other_agents_observed
        assert 0 <= tagging_distance <= 1
        self.distance_margin_for_reward = f32(tagging_distance * self.grid_length)
        assert tag_reward_for_tagger >= 0 and step_penalty_for_tagger <= 0
        assert tag_penalty_for_runner <= 0 and step_reward_for_runner >= 0
        assert end_of_game_reward_for_runner >= 0
        self.tag_reward_for_tagger = f32(tag_reward_for_tagger)
        self.step_penalty_for_tagger = f32(step_penalty_for_tagger)
        self.tag_penalty_for_runner = f32(tag_penalty_for_runner)
        self.step_reward_for_runner = f32(step_reward_for_runner)
        self.end_of_game_reward_for_runner = f32(end_of_game_reward_for_runner)
        self.step_rewards = [
            self.step_penalty_for_tagger if is_tagger[a] else self.step_reward_for_runner
            for a in range(N)]
        self._types = types.astype(np.int32)
        self.edge_hit_reward_penalty = None
        self.still_in_the_game = None
        self.env_backend = env_backend
        # the reference env allocates 2 x [E, N, N-1] global scratch for its kNN sort;
        # the B200 kernel keeps the sweep on chip, so by default they are 1-element stubs
        # The reference always allocates two [E, N, N-1] scratch arrays for the neighbour
        # search; here they are only needed when the exact-tie path's per-warp lists
        # (8 N bytes per warp) do not fit next to the staged state in shared memory (>~400
        # agents), or when the observation dimension K + 2 exceeds the sorting network
        self._scratch_requested = allocate_reference_scratch
        self.allocate_reference_scratch = allocate_reference_scratch
        self.runners_at_reset = dict(self.runners)

    def seed(self, seed=None):
        self.np_random.seed(seed)
        return [seed]

    # ------------------------------------------------------------------ state
    def set_global_state(self, key=None, value=None, t=None, dtype=None):
        assert key is not None
        if key not in self.global_state:
            self.global_state[key] = np.zeros(
                (self.episode_length + 1, self.num_agents),
                dtype=self.float_dtype if dtype is None else dtype)
        if t is not None and value is not None:
            assert isinstance(value, np.ndarray)
            assert value.shape[0] == self.global_state[key].shape[1]
            self.global_state[key][t] = value

    def reset(self):
        self.timestep = 0
        self.global_state = {}
        for key, val in zip(_STATE_KEYS, (
                self.starting_location_x, self.starting_location_y, self.starting_speeds,
                self.starting_directions, self.starting_accelerations)):
            self.set_global_state(key=key, value=np.asarray(val), t=0)
        self.still_in_the_game = np.ones(self.num_agents, dtype=self.int_dtype)
        self.global_state["still_in_the_game"] = np.ones(
            (self.episode_length + 1, self.num_agents), dtype=self.int_dtype)
        self.edge_hit_reward_penalty = np.zeros(self.num_agents, dtype=self.float_dtype)
        self.runners = dict(self.runners_at_reset)
        self.num_runners = len(self.runners)
        return self.generate_observation()

    # ------------------------------------------------------------------ CPU step
    def update_state(self, delta_accelerations, delta_turns):
        f32 = np.float32
        t = self.timestep
        gs = self.global_state
        alive = self.still_in_the_game.astype(f32)
        two_pi = f32(6.283185308)
        direction = np.fmod(gs["direction"][t - 1] + delta_turns.astype(f32), two_pi) * alive
        direction = np.where(direction < 0, direction + two_pi, direction).astype(f32)
        acc = (gs["acceleration"][t - 1] + delta_accelerations.astype(f32)).astype(f32)
        cap = (self.max_speed * np.asarray(self.skill_levels, dtype=f32)).astype(f32)
        speed = (np.clip(gs["speed"][t - 1] + acc, f32(0), cap) * alive).astype(f32)
        acc = np.where((speed <= 0) | (speed >= cap), f32(0), acc).astype(f32)
        x = (gs["loc_x"][t - 1] + speed * np.cos(direction)).astype(f32)
        y = (gs["loc_y"][t - 1] + speed * np.sin(direction)).astype(f32)
        L = self.grid_length
        crossed = (x < 0) | (x > L) | (y < 0) | (y > L)
        self.edge_hit_reward_penalty = np.where(crossed, self.edge_hit_penalty, f32(0)).astype(f32)
        for key, val in zip(_STATE_KEYS, (np.clip(x, f32(0), L), np.clip(y, f32(0), L),
                                          speed, direction, acc)):
            self.set_global_state(key=key, value=val.astype(f32), t=t)

    def _pairwise_distance(self):
        t = self.timestep
        x = self.global_state["loc_x"][t].astype(np.float64)
        y = self.global_state["loc_y"][t].astype(np.float64)
        dx = (x[:, None].astype(np.float32) - x[None, :].astype(np.float32)).astype(np.float64)
        dy = (y[:, None].astype(np.float32) - y[None, :].astype(np.float32)).astype(np.float64)
        return np.sqrt(dx * dx + dy * dy).astype(np.float32)

    def k_nearest_neighbors(self, agent_id, k, dist=None):
        if dist is None:
            dist = self._pairwise_distance()
        cand = np.array([b for b in range(self.num_agents)
                         if b != agent_id and self.still_in_the_game[b]], dtype=np.int64)
        if len(cand) == 0:
            return []
        order = np.argsort(dist[agent_id, cand], kind="stable")[:k]
        return cand[order].tolist()

    def generate_observation(self):
        f32 = np.float32
        t, N, gs = self.timestep, self.num_agents, self.global_state
        x, y, sp = gs["loc_x"][t], gs["loc_y"][t], gs["speed"][t]
        acc, di = gs["acceleration"][t], gs["direction"][t]
        alive = self.still_in_the_game
        diag = np.sqrt(2.0) * np.float64(self.grid_length)
        vnorm = f32(self.max_speed + self.eps)
        two_pi = f32(6.283185308)
        time = f32(f32(t) / f32(self.episode_length))

        def features(a, others):
            o = np.asarray(others, dtype=np.int64)
            return [
                ((x[o] - x[a]).astype(np.float64) / diag).astype(f32),
                ((y[o] - y[a]).astype(np.float64) / diag).astype(f32),
                ((sp[o] - sp[a]) / vnorm).astype(f32),
                ((acc[o] - acc[a]) / vnorm).astype(f32),
                ((di[o] - di[a]) / two_pi).astype(f32),
            ]

        obs = {}
        if self.use_full_observation:
            M = N - 1
            for a in range(N):
                others = [b for b in range(N) if b != a]
                row = np.zeros(7 * M + 1, dtype=f32)
                row[5 * M:6 * M] = self._types[others]
                row[6 * M:7 * M] = alive[others]
                if alive[a]:
                    row[:5 * M] = np.concatenate(features(a, others))
                    row[7 * M] = time
                obs[a] = row
            return obs
        K = self.num_other_agents_observed
        dist = self._pairwise_distance()
        for a in range(N):
            row = np.zeros(7 * K + 1, dtype=f32)
            if alive[a]:
                nn = self.k_nearest_neighbors(a, K, dist)
                k = len(nn)
                if k:
                    feats = features(a, nn)
                    for f in range(5):
                        row[f * K:f * K + k] = feats[f]
                    row[5 * K:5 * K + k] = self._types[nn]
                    row[6 * K:6 * K + k] = alive[nn]
                row[7 * K] = time
            obs[a] = row
        return obs

    def compute_reward(self):
        f32 = np.float32
        t, N = self.timestep, self.num_agents
        alive_at_entry = self.still_in_the_game.copy()
        rew = np.zeros(N, dtype=f32)
        step_rewards = np.asarray(self.step_rewards, dtype=f32)
        rew = np.where(alive_at_entry > 0, self.edge_hit_reward_penalty + step_rewards, f32(0)).astype(f32)
        taggers = sorted(self.taggers)
        dist = self._pairwise_distance()
        for r in sorted(self.runners):
            d = dist[r, taggers]
            j = int(np.argmin(d))          # first minimum == strict-< scan in id order
            if d[j] < self.distance_margin_for_reward:
                rew[r] += self.tag_penalty_for_runner
                rew[taggers[j]] += self.tag_reward_for_tagger
                if self.runner_exits_game_after_tagged:
                    self.still_in_the_game[r] = 0
                    del self.runners[r]
                    self.num_runners -= 1
                    self.global_state["still_in_the_game"][t:, r] = 0
        if t == self.episode_length:
            for r in self.runners:
                rew[r] += self.end_of_game_reward_for_runner
        return {a: rew[a] for a in range(N)}

    # ------------------------------------------------------------------ device data
    def get_data_dictionary(self):
        N, K = self.num_agents, self.num_other_agents_observed
        d = DataFeed()
        for key in _STATE_KEYS:
            d.add_data(name=key, data=self.global_state[key][0],
                       save_copy_and_apply_at_reset=True)
        d.add_data(name="agent_types", data=[self.agent_type[a] for a in range(N)])
        d.add_data(name="num_runners", data=self.num_runners,
                   save_copy_and_apply_at_reset=True)
        d.add_data(name="num_other_agents_observed", data=K)
        d.add_data(name="grid_length", data=self.grid_length)
        d.add_data(name="edge_hit_reward_penalty", data=self.edge_hit_reward_penalty,
                   save_copy_and_apply_at_reset=True)
        d.add_data(name="step_rewards", data=self.step_rewards)
        d.add_data(name="edge_hit_penalty", data=self.edge_hit_penalty)
        d.add_data(name="max_speed", data=self.max_speed)
        d.add_data(name="acceleration_actions", data=self.acceleration_actions)
        d.add_data(name="turn_actions", data=self.turn_actions)
        d.add_data(name="skill_levels", data=self.skill_levels)
        d.add_data(name="use_full_observation", data=self.use_full_observation)
        d.add_data(name="distance_margin_for_reward", data=self.distance_margin_for_reward)
        d.add_data(name="tag_reward_for_tagger", data=self.tag_reward_for_tagger)
        d.add_data(name="tag_penalty_for_runner", data=self.tag_penalty_for_runner)
        d.add_data(name="end_of_game_reward_for_runner",
                   data=self.end_of_game_reward_for_runner)
        n_warps = (N + 31) // 32
        bpe = int(getattr(getattr(self, "cuda_function_manager", None), "blocks_per_env", 1) or 1)
        # the single-CTA kernel spills its exact-path lists to the reference's [N, N-1] scratch
        # for large envs; the cluster kernel (blocks_per_env > 1) keeps everything on chip
        # and so does the one-CTA-per-env variant of that kernel that serves N > 320
        packed = N <= 320 or K + 2 > 16 or bool(self.use_full_observation)
        if bpe <= 1 and packed and (36 * N + 8 * N * n_warps > 60000 or K + 2 > 16):
            self.allocate_reference_scratch = True
        if self.allocate_reference_scratch:
            d.add_data(name="neighbor_distances",
                       data=np.zeros((N, N - 1), dtype=np.float32),
                       save_copy_and_apply_at_reset=True)
            d.add_data(name="neighbor_ids_sorted_by_distance",
                       data=np.zeros((N, N - 1), dtype=np.int32),
                       save_copy_and_apply_at_reset=True)
        else:
            d.add_data(name="neighbor_distances", data=np.zeros(1, dtype=np.float32))
            d.add_data(name="neighbor_ids_sorted_by_distance",
                       data=np.zeros(1, dtype=np.int32))
        d.add_data(name="nearest_neighbor_ids", data=np.zeros((N, K), dtype=np.int32),
                   save_copy_and_apply_at_reset=True)
        d.add_data(name="runner_exits_game_after_tagged",
                   data=self.runner_exits_game_after_tagged)
        d.add_data(name="still_in_the_game", data=self.still_in_the_game,
                   save_copy_and_apply_at_reset=True)
        return d

    # ------------------------------------------------------------------ step
    def step_env_range(self, e0, e1):
        """Device step of the env replicas [e0, e1) only, on the current stream (replicas are
        independent; the caller accounts for `timestep`).  Used by
        EnvWrapper.step_with_host_buffers to overlap the host copies of one group of envs with
        the step of the next."""
        dm = self.cuda_data_manager
        n_all = int(dm.meta_info("n_envs"))
        assert 0 <= e0 < e1 <= n_all and self.env_backend != "cpu"
        args = []
        for arg in _STEP_ARGS:
            if isinstance(arg, tuple):
                args.append(dm.meta_info(arg[0]))
            elif arg in ("neighbor_distances", "neighbor_ids_sorted_by_distance") and \
                    not self.allocate_reference_scratch:
                args.append(None)
            elif arg in _PER_ENV_ARGS:
                args.append(dm.data_on_device_via_torch(arg)[e0:e1])
            else:
                args.append(dm.device_data(arg))
        self.cuda_step(*args, block=self.cuda_function_manager.block,
                       grid=self.cuda_function_manager.grid, n_envs=e1 - e0)

    def step(self, actions=None):
        self.timestep += 1
        if self.env_backend != "cpu":
            args = list(self.cuda_step_function_feed(_STEP_ARGS))
            if not self.allocate_reference_scratch:
                args[19] = None   # neighbor_distances: kept on chip
                args[20] = None   # neighbor_ids_sorted_by_distance
            if self.env_backend == "numba":
                self.cuda_step[self.cuda_function_manager.grid,
                               self.cuda_function_manager.block](*args)
            else:
                self.cuda_step(*args, block=self.cuda_function_manager.block,
                               grid=self.cuda_function_manager.grid)
            return None
        assert isinstance(actions, dict) and len(actions) == self.num_agents
        a0 = np.array([actions[a][0] for a in range(self.num_agents)])
        a1 = np.array([actions[a][1] for a in range(self.num_agents)])
        assert ((0 <= a0) & (a0 <= self.num_acceleration_levels)).all()
        assert ((0 <= a1) & (a1 <= self.num_turn_levels)).all()
        self.update_state(self.acceleration_actions[a0], self.turn_actions[a1])
        obs = self.generate_observation()
        rew = self.compute_reward()
        done = {"__all__": (self.timestep >= self.episode_length) or (self.num_runners == 0)}
        return obs, rew, done, {}
