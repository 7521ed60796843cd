"""Generate the committed golden fixtures from the REAL reference.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py

What it writes (all under tests/golden/):
  gridworld_cuda_golden.npz   the literal known-answer vectors held by the
      reference's own CUDA test (tests/example_envs/pycuda_tests/
      test_tag_gridworld_step_cuda.py:132-708), extracted from its AST -- no
      reference code is executed or copied for this file.
  gridworld_numpy_<cfg>.npz   trajectories of the reference NumPy TagGridWorld
      (example_envs/tag_gridworld/tag_gridworld.py) driven by seeded random actions.
  tag_continuous_numpy_<cfg>.npz  trajectories of the reference NumPy TagContinuous
      (example_envs/tag_continuous/tag_continuous.py:796-887) for the four configs of
      the reference's own consistency test (tests/example_envs/pycuda_tests/
      test_tag_continuous.py:15-80) plus two partial-observation configs shaped like
      BASELINE.json config 2.

`gym` is not installed here; a minimal `gym.spaces` stand-in (Box / Discrete /
MultiDiscrete / Dict) is injected into sys.modules so the reference imports.
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def install_gym_shim():
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Space:
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.dtype = low, high, dtype
            self.shape = tuple(shape) if shape is not None else np.shape(low)

    class Discrete(Space):
        def __init__(self, n):
            self.n = int(n)

    class MultiDiscrete(Space):
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)

    class Dict(Space, dict):
        def __init__(self, spaces_dict=None):
            dict.__init__(self, spaces_dict or {})
            self.spaces = self

    for cls in (Space, Box, Discrete, MultiDiscrete, Dict):
        setattr(spaces, cls.__name__, cls)
    gym.spaces = spaces
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces


# --------------------------------------------------------------------------- #
# 1. literal golden vectors of the reference's CUDA gridworld test (AST only)
# --------------------------------------------------------------------------- #
def _literal_arrays(func_node, names):
    """Collect `name = np.array(<literal>)` assignments, in order of appearance."""
    found = {n: [] for n in names}
    for node in ast.walk(func_node):
        if isinstance(node, ast.Assign) and len(node.targets) == 1:
            tgt = node.targets[0]
            if isinstance(tgt, ast.Name) and tgt.id in found:
                call = node.value
                if isinstance(call, ast.Call) and call.args:
                    try:
                        lit = ast.literal_eval(call.args[0])
                    except ValueError:
                        continue  # e.g. torch.from_numpy(name): not a literal
                    found[tgt.id].append((node.lineno, np.array(lit)))
    return found


def extract_gridworld_cuda_golden():
    path = os.path.join(
        REF, "tests/example_envs/pycuda_tests/test_tag_gridworld_step_cuda.py"
    )
    tree = ast.parse(open(path).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "int_step":
            arrs = _literal_arrays(node, [])
            for sub in ast.walk(node):
                if isinstance(sub, ast.Call) and getattr(sub.func, "attr", "") == "add_data":
                    kw = {k.arg: k.value for k in sub.keywords}
                    name = ast.literal_eval(kw["name"]) if isinstance(kw["name"], ast.Constant) else None
                    if name in ("loc_x", "loc_y"):
                        out[f"init_{name}"] = np.array(
                            ast.literal_eval(kw["data"].args[0]), dtype=np.int32
                        )
            for sub in ast.walk(node):
                if isinstance(sub, ast.Assign) and isinstance(sub.targets[0], ast.Name):
                    nm = sub.targets[0].id
                    if nm in (
                        "wall_hit_penalty", "tag_reward_for_tagger",
                        "tag_penalty_for_runner", "step_cost_for_tagger",
                        "world_boundary", "kIndexToActionArr",
                    ):
                        out[nm] = np.array(ast.literal_eval(sub.value))
        if isinstance(node, ast.FunctionDef) and node.name == "test_step":
            names = [
                "agent_distribution", "ref_rewards", "ref_observations",
                "ref_actions", "ref_done",
            ]
            arrs = _literal_arrays(node, names)
            for nm in names:
                vals = sorted(arrs[nm])
                for i, (_, a) in enumerate(vals):
                    out[f"{nm}_step{i + 1}"] = a
    assert out["ref_observations_step1"].shape == (2, 5, 21)
    assert out["ref_observations_step2"].shape == (2, 5, 21)
    np.savez_compressed(os.path.join(HERE, "gridworld_cuda_golden.npz"), **out)
    return out


# --------------------------------------------------------------------------- #
# 2. reference NumPy env trajectories
# --------------------------------------------------------------------------- #
TC_TEST_CONFIGS = {
    # tests/example_envs/pycuda_tests/test_tag_continuous.py:15-80
    "test1": dict(num_taggers=2, num_runners=3, max_acceleration=1, max_turn=np.pi / 4,
                  num_acceleration_levels=3, num_turn_levels=3, grid_length=10,
                  episode_length=100, seed=274880, skill_level_runner=1,
                  skill_level_tagger=1, use_full_observation=True,
                  runner_exits_game_after_tagged=True, tagging_distance=0.0),
    "test2": dict(num_taggers=4, num_runners=1, max_acceleration=0.05, max_turn=np.pi / 4,
                  num_acceleration_levels=3, num_turn_levels=3, grid_length=10,
                  episode_length=100, step_penalty_for_tagger=-0.1, seed=428096,
                  skill_level_runner=1, skill_level_tagger=2, use_full_observation=False,
                  runner_exits_game_after_tagged=False, tagging_distance=0.25),
    "test3": dict(num_taggers=1, num_runners=4, max_acceleration=2, max_turn=np.pi / 2,
                  num_acceleration_levels=3, num_turn_levels=3, grid_length=10,
                  episode_length=100, step_reward_for_runner=0.1, seed=654208,
                  skill_level_runner=1, skill_level_tagger=0.5, use_full_observation=False,
                  runner_exits_game_after_tagged=True),
    "test4": dict(num_taggers=3, num_runners=2, max_acceleration=0.05, max_turn=np.pi,
                  num_acceleration_levels=3, num_turn_levels=3, grid_length=10,
                  episode_length=100, seed=121024, skill_level_runner=0.5,
                  skill_level_tagger=1, use_full_observation=True,
                  runner_exits_game_after_tagged=False),
    # shaped like warp_drive/training/run_configs/tag_continuous.yaml:10-34, smaller
    "partial_mid": dict(num_taggers=3, num_runners=20, grid_length=10.0, episode_length=60,
                        max_acceleration=0.1, min_acceleration=-0.1, max_turn=2.356,
                        min_turn=-2.356, num_acceleration_levels=20, num_turn_levels=20,
                        skill_level_runner=1.0, skill_level_tagger=1.0, max_speed=1.0,
                        seed=274880, use_full_observation=False,
                        runner_exits_game_after_tagged=True, num_other_agents_observed=5,
                        tag_reward_for_tagger=10.0, tag_penalty_for_runner=-10.0,
                        edge_hit_penalty=-0.5, end_of_game_reward_for_runner=1.0,
                        tagging_distance=0.05),
    # BASELINE.json config 2 env (5 taggers + 100 runners, K=10), short horizon
    "config2_short": dict(num_taggers=5, num_runners=100, grid_length=20.0,
                          episode_length=500, max_acceleration=0.1, min_acceleration=-0.1,
                          max_turn=2.356, min_turn=-2.356, num_acceleration_levels=20,
                          num_turn_levels=20, skill_level_runner=1.0,
                          skill_level_tagger=1.0, max_speed=1.0, seed=274880,
                          use_full_observation=False, runner_exits_game_after_tagged=True,
                          num_other_agents_observed=10, tag_reward_for_tagger=10.0,
                          tag_penalty_for_runner=-10.0, step_penalty_for_tagger=0.0,
                          step_reward_for_runner=0.0, edge_hit_penalty=0.0,
                          end_of_game_reward_for_runner=1.0, tagging_distance=0.02),
}
TC_STEPS = {"test1": 100, "test2": 100, "test3": 100, "test4": 100,
            "partial_mid": 60, "config2_short": 25}


def _cfg_to_saveable(cfg):
    return {f"cfg__{k}": np.asarray(v) for k, v in cfg.items()}


def record_tag_continuous(name, cfg, n_steps, action_seed=1234):
    from example_envs.tag_continuous.tag_continuous import TagContinuous

    env = TagContinuous(**cfg)
    obs0 = env.reset()
    N = env.num_agents
    rs = np.random.RandomState(action_seed)
    na, nt = len(env.acceleration_actions), len(env.turn_actions)
    dd = env.get_data_dictionary()
    out = _cfg_to_saveable(cfg)
    for key in ("loc_x", "loc_y", "speed", "direction", "acceleration", "agent_types",
                "step_rewards", "acceleration_actions", "turn_actions", "skill_levels",
                "still_in_the_game", "edge_hit_reward_penalty"):
        out[f"init__{key}"] = np.array(dd[key]["data"], copy=True)
    for key in ("num_runners", "grid_length", "edge_hit_penalty", "max_speed",
                "distance_margin_for_reward", "tag_reward_for_tagger",
                "tag_penalty_for_runner", "end_of_game_reward_for_runner",
                "num_other_agents_observed", "use_full_observation",
                "runner_exits_game_after_tagged"):
        out[f"init__{key}"] = np.array(dd[key]["data"], copy=True)
    out["episode_length"] = np.asarray(env.episode_length)
    out["obs0"] = np.stack([np.asarray(obs0[a], dtype=np.float64) for a in range(N)])
    acts, obs, rews, dones, alive, lx, ly, sp, di, ac = ([] for _ in range(10))
    for _ in range(n_steps):
        a = np.stack([rs.randint(0, na, N), rs.randint(0, nt, N)], axis=1).astype(np.int32)
        o, r, d, _ = env.step({i: a[i] for i in range(N)})
        acts.append(a)
        obs.append(np.stack([np.asarray(o[i], dtype=np.float64) for i in range(N)]))
        rews.append(np.array([r[i] for i in range(N)], dtype=np.float64))
        dones.append(bool(d["__all__"]))
        alive.append(env.still_in_the_game.copy())
        t = env.timestep
        lx.append(env.global_state["loc_x"][t].copy())
        ly.append(env.global_state["loc_y"][t].copy())
        sp.append(env.global_state["speed"][t].copy())
        di.append(env.global_state["direction"][t].copy())
        ac.append(env.global_state["acceleration"][t].copy())
        if d["__all__"]:
            break
    out.update(actions=np.stack(acts), obs=np.stack(obs).astype(np.float32),
               rewards=np.stack(rews).astype(np.float32), done=np.array(dones),
               still_in_the_game=np.stack(alive).astype(np.int32),
               loc_x=np.stack(lx), loc_y=np.stack(ly), speed=np.stack(sp),
               direction=np.stack(di), acceleration=np.stack(ac))
    np.savez_compressed(os.path.join(HERE, f"tag_continuous_numpy_{name}.npz"), **out)
    return out


GW_CONFIGS = {
    # tests/example_envs/pycuda_tests/test_tag_gridworld.py:14-37
    "test1": dict(num_taggers=4, grid_length=4, episode_length=20, seed=27,
                  wall_hit_penalty=0.1, tag_reward_for_tagger=10.0,
                  tag_penalty_for_runner=2.0, step_cost_for_tagger=0.01,
                  use_full_observation=True),
    "test2": dict(num_taggers=4, grid_length=4, episode_length=20, seed=27,
                  wall_hit_penalty=0.1, tag_reward_for_tagger=10.0,
                  tag_penalty_for_runner=2.0, step_cost_for_tagger=0.01,
                  use_full_observation=False),
    # BASELINE.json config 1
    "config1": dict(num_taggers=4, grid_length=10, episode_length=100, seed=20,
                    wall_hit_penalty=0.1, tag_reward_for_tagger=10.0,
                    tag_penalty_for_runner=5.0, step_cost_for_tagger=0.01),
}


def record_tag_gridworld(name, cfg, n_episodes=3, action_seed=0):
    from example_envs.tag_gridworld.tag_gridworld import TagGridWorld

    env = TagGridWorld(**cfg)
    N = env.num_agents
    rs = np.random.RandomState(action_seed)
    out = _cfg_to_saveable(cfg)
    out["init__loc_x"] = np.asarray(env.starting_location_x, dtype=np.int32)
    out["init__loc_y"] = np.asarray(env.starting_location_y, dtype=np.int32)
    acts, obs, rews, dones, lx, ly, ep = ([] for _ in range(7))
    for e in range(n_episodes):
        o0 = env.reset()
        if e == 0:
            out["obs0"] = np.stack([np.asarray(o0[a], dtype=np.float64) for a in range(N)]).astype(np.float32)
        while True:
            a = rs.randint(0, 5, N).astype(np.int32)
            o, r, d, _ = env.step({i: int(a[i]) for i in range(N)})
            acts.append(a)
            obs.append(np.stack([np.asarray(o[i], dtype=np.float64) for i in range(N)]))
            rews.append(np.array([r[i] for i in range(N)], dtype=np.float64))
            dones.append(bool(d["__all__"]))
            lx.append(env.global_state["loc_x"][env.timestep].copy())
            ly.append(env.global_state["loc_y"][env.timestep].copy())
            ep.append(e)
            if d["__all__"]:
                break
    out.update(actions=np.stack(acts), obs=np.stack(obs).astype(np.float32),
               rewards=np.stack(rews).astype(np.float32), done=np.array(dones),
               loc_x=np.stack(lx).astype(np.int32), loc_y=np.stack(ly).astype(np.int32),
               episode=np.array(ep))
    np.savez_compressed(os.path.join(HERE, f"gridworld_numpy_{name}.npz"), **out)
    return out


def main():
    assert os.path.isdir(REF), "the reference tree is only present in the build container"
    install_gym_shim()
    sys.path.insert(0, REF)
    g = extract_gridworld_cuda_golden()
    print("gridworld_cuda_golden:", sorted(g))
    for name, cfg in GW_CONFIGS.items():
        o = record_tag_gridworld(name, cfg)
        print("gridworld", name, o["obs"].shape, int(o["done"].sum()), "episodes ended")
    for name, cfg in TC_TEST_CONFIGS.items():
        o = record_tag_continuous(name, cfg, TC_STEPS[name])
        print("tag_continuous", name, o["obs"].shape, "alive at end", int(o["still_in_the_game"][-1].sum()))


if __name__ == "__main__":
    main()
