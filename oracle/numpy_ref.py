"""NumPy restatement of the reference's *Python CPU* TagContinuous step -- the reported
CPU baseline.  TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

BASELINE.json's north_star asks for "the reference's own Python/NumPy CPU step() timed on
the same box's host cores".  /root/reference does not exist on the GPU box, so the
algorithm of example_envs/tag_continuous/tag_continuous.py is restated here with the same
structure and therefore the same cost profile: per-agent Python loops, a heap-based
k-nearest search (heapq.nsmallest over per-pair np.sqrt calls) and NumPy vector math for
the kinematics.  Each method cites the reference lines it follows.  Pinned against
trajectories recorded from the real reference (tests/golden/tag_continuous_numpy_*.npz,
tests/test_oracle_cpu.py::test_numpy_port_matches_reference).
"""
import heapq

import numpy as np


class TagContinuousNumpyRef:
    """One env replica.  `cfg` and `init` use the names of the reference's
    get_data_dictionary (tag_continuous.py:680-756)."""

    def __init__(self, cfg, init):
        self.cfg = cfg
        self.N = len(init["loc_x"])
        self.types = np.asarray(cfg["agent_types"])
        self.init = {k: np.array(v, copy=True) for k, v in init.items()}
        self.f32 = np.float32
        self.grid_length = np.float32(cfg["grid_length"])
        self.grid_diagonal = self.grid_length * np.sqrt(2)           # :147
        self.max_speed = np.float32(cfg["max_speed"])
        self.eps = np.float32(1e-10)                                  # :132
        self.skill_levels = [np.float32(s) for s in cfg["skill_levels"]]
        self.step_rewards = [np.float32(s) for s in cfg["step_rewards"]]
        self.episode_length = int(cfg["episode_length"])
        self.K = int(cfg["num_other_agents_observed"])
        self.reset()

    def reset(self):                                                  # :758-794
        self.timestep = 0
        self.state = {k: np.array(self.init[k], dtype=np.float32, copy=True)
                      for k in ("loc_x", "loc_y", "speed", "direction", "acceleration")}
        self.still_in_the_game = np.ones(self.N, dtype=np.int32)
        self.edge_hit_reward_penalty = np.zeros(self.N, dtype=np.float32)
        self.taggers = [a for a in range(self.N) if self.types[a] == 1]
        self.runners = {a: True for a in range(self.N) if self.types[a] == 0}
        self.num_runners = len(self.runners)

    def update_state(self, delta_accelerations, delta_turns):         # :339-401
        s, f32 = self.state, self.f32
        direction = ((s["direction"] + delta_turns) % (2 * np.pi)
                     * self.still_in_the_game).astype(f32)
        acc = s["acceleration"] + delta_accelerations
        max_speed = self.max_speed * np.array(self.skill_levels)
        speed = f32(np.clip(s["speed"] + acc, 0.0, max_speed) * self.still_in_the_game)
        acc = acc * (speed > 0) * (speed < max_speed)
        x = f32(s["loc_x"] + speed * np.cos(direction))
        y = f32(s["loc_y"] + speed * np.sin(direction))
        crossed = ~((x >= 0) & (x <= self.grid_length) & (y >= 0) & (y <= self.grid_length))
        self.edge_hit_reward_penalty = np.float32(self.cfg["edge_hit_penalty"]) * crossed
        s["loc_x"] = f32(np.clip(x, 0.0, self.grid_length))
        s["loc_y"] = f32(np.clip(y, 0.0, self.grid_length))
        s["speed"], s["direction"], s["acceleration"] = speed, direction, f32(acc)

    def compute_distance(self, a1, a2):                               # :403-420
        s = self.state
        return np.sqrt((s["loc_x"][a1] - s["loc_x"][a2]) ** 2
                       + (s["loc_y"][a1] - s["loc_y"][a2]) ** 2).astype(self.f32)

    def k_nearest_neighbors(self, agent_id, k):                       # :422-444
        pairs = []
        for ag in range(self.N):
            if ag != agent_id and self.still_in_the_game[ag]:
                pairs.append((ag, self.compute_distance(agent_id, ag)))
        return [p[0] for p in heapq.nsmallest(k, pairs, key=lambda x: x[1])][: self.K]

    def generate_observation(self):                                   # :446-610
        s = self.state
        norm = None
        for key, scale in (("loc_x", self.grid_diagonal), ("loc_y", self.grid_diagonal),
                           ("speed", self.max_speed + self.eps),
                           ("acceleration", self.max_speed + self.eps),
                           ("direction", 2 * np.pi)):
            row = s[key] / scale
            norm = row if norm is None else np.vstack((norm, row))
        types = np.array(self.types)
        time = np.array([float(self.timestep) / self.episode_length])
        obs = {}
        if int(self.cfg["use_full_observation"]):
            for a in range(self.N):
                others = [i for i in range(self.N) if i != a]
                if self.still_in_the_game[a]:
                    block = np.vstack((norm - norm[:, a].reshape(-1, 1), types,
                                       self.still_in_the_game))[:, others].reshape(-1)
                    obs[a] = np.concatenate([block, time])
                else:
                    block = np.vstack((np.zeros_like(norm), types,
                                       self.still_in_the_game))[:, others].reshape(-1)
                    obs[a] = np.concatenate([block, np.array([0.0])])
            return obs
        K = self.K
        zero = np.zeros(7 * K + 1)
        for a in range(self.N):
            obs[a] = zero
            if not self.still_in_the_game[a]:
                continue
            nn = self.k_nearest_neighbors(a, k=K)
            pad = K - len(nn)
            g = np.hstack((norm[:, nn] - norm[:, a].reshape(-1, 1), np.zeros((5, pad))))
            t = np.hstack((types[nn], np.zeros(pad)))
            al = np.hstack((self.still_in_the_game[nn], np.zeros(pad)))
            obs[a] = np.concatenate([np.vstack((g, t, al)).reshape(-1), time])
        return obs

    def compute_reward(self):                                         # :612-678
        s = self.state
        rew = {a: 0.0 for a in range(self.N)}
        taggers = sorted(self.taggers)
        runners = sorted(self.runners)
        if self.num_runners > 0:
            rx, ry = s["loc_x"][runners], s["loc_y"][runners]
            tx, ty = s["loc_x"][taggers], s["loc_y"][taggers]
            nt, nr = len(taggers), self.num_runners
            d = np.sqrt((np.repeat(rx, nt) - np.tile(tx, nr)) ** 2
                        + (np.repeat(ry, nt) - np.tile(ty, nr)) ** 2).reshape(nr, nt)
            dmin, amin = np.min(d, axis=1), np.argmin(d, axis=1)
            nearest = [taggers[i] for i in amin]
        for a in range(self.N):
            if self.still_in_the_game[a]:
                rew[a] += self.edge_hit_reward_penalty[a]
                rew[a] += self.step_rewards[a]
        margin = np.float32(self.cfg["distance_margin_for_reward"])
        for idx, r in enumerate(runners):
            if dmin[idx] < margin:
                rew[r] += np.float32(self.cfg["tag_penalty_for_runner"])
                rew[nearest[idx]] += np.float32(self.cfg["tag_reward_for_tagger"])
                if int(self.cfg["runner_exits_game_after_tagged"]):
                    self.still_in_the_game[r] = 0
                    del self.runners[r]
                    self.num_runners -= 1
        if self.timestep == self.episode_length:
            for r in self.runners:
                rew[r] += np.float32(self.cfg["end_of_game_reward_for_runner"])
        return rew

    def step(self, actions):                                          # :796-887 (cpu branch)
        """actions: int array [N, 2]"""
        self.timestep += 1
        d_acc = np.asarray(self.cfg["acceleration_actions"])[actions[:, 0]]
        d_turn = np.asarray(self.cfg["turn_actions"])[actions[:, 1]]
        self.update_state(d_acc, d_turn)
        obs = self.generate_observation()
        rew = self.compute_reward()
        done = (self.timestep >= self.episode_length) or (self.num_runners == 0)
        return obs, rew, done


def _worker(args):
    """Time `n_steps` env steps of one replica on one core (bench.py cpu_baseline)."""
    import time

    cfg, init, n_steps, warmup, seed = args
    env = TagContinuousNumpyRef(cfg, init)
    rs = np.random.RandomState(seed)
    na, nt = len(cfg["acceleration_actions"]), len(cfg["turn_actions"])
    acts = np.stack([rs.randint(0, na, (n_steps + warmup, env.N)),
                     rs.randint(0, nt, (n_steps + warmup, env.N))], axis=-1)
    for t in range(warmup):
        env.step(acts[t])
    t0 = time.perf_counter()
    for t in range(warmup, warmup + n_steps):
        _, _, done = env.step(acts[t])
        if done:
            env.reset()
    return time.perf_counter() - t0, n_steps * env.N


def timed_agent_steps_per_sec(cfg, init, n_steps, warmup=2, n_procs=None):
    """Aggregate agent-steps/s over `n_procs` processes (one replica each)."""
    import multiprocessing as mp
    import os
    import time

    n_procs = n_procs or os.cpu_count() or 1
    jobs = [(cfg, init, n_steps, warmup, 1000 + i) for i in range(n_procs)]
    t0 = time.perf_counter()
    if n_procs == 1:
        res = [_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(n_procs) as pool:
            res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    work = sum(r[1] for r in res)
    slowest = max(r[0] for r in res)
    return {"agent_steps_per_sec": work / slowest, "wall_s": wall, "cores": n_procs,
            "steps_per_proc": n_steps, "n_agents": len(init["loc_x"])}
