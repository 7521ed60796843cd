#!/usr/bin/env python
"""Rollout step of the GENERIC multi-launch path (no env-specific fused step kernel):
tag_gridworld 2000 envs x 5 agents ([32, 32] policy, T = 20 steps per CUDA graph), ms per
timestep and libwdb200 launches per timestep; the same engine with the torch-op bookkeeping
(RolloutEngine.bookkeep_torch: what the native wdb_rollout_bookkeep launch replaced) beside it.
Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import yaml  # noqa: E402

from warp_drive_b200 import lib as wlib  # noqa: E402


def build(E, T, torch_bookkeep):
    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.tag_gridworld import CUDATagGridWorld
    from warp_drive_b200.managers.function_manager import CUDASampler
    from warp_drive_b200.training.models.fully_connected import FullyConnected
    from warp_drive_b200.training.rollout import RolloutEngine
    from warp_drive_b200.training.utils.data_loader import create_and_push_data_placeholders

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "warp_drive_b200", "training", "run_configs",
                           "tag_gridworld.yaml"), encoding="utf8") as fp:
        cfg = yaml.safe_load(fp)
    env = CUDATagGridWorld(**cfg["env"])
    w = EnvWrapper(env, num_envs=E, env_backend="b200")
    w.reset_all_envs()
    pm = {"shared": list(range(env.num_agents))}
    s = CUDASampler(w.cuda_function_manager)
    create_and_push_data_placeholders(env_wrapper=w, action_sampler=s,
                                      policy_tag_to_agent_id_map=pm,
                                      training_batch_size_per_env=T)
    s.init_random(7)
    torch.manual_seed(0)
    models = {"shared": FullyConnected(w, cfg["policy"]["shared"]["model"], "shared", pm).cuda().eval()}
    eng = RolloutEngine(w, models, pm, s, T, use_cuda_graph=True)
    assert eng.fused is None and eng.sa is None
    if torch_bookkeep:
        eng.bookkeep = eng.bookkeep_torch
    return w, eng


def time_engine(eng, T, reps=20, rounds=5):
    for _ in range(3):
        eng.rollout()
    torch.cuda.synchronize()
    out = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            eng.rollout()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / (reps * T))
    return sorted(out)[len(out) // 2]


def main():
    E, T = 2000, 20
    res = {"workload": f"tag_gridworld {E} envs x 5 agents, generic rollout path, T = {T}"}
    for name, tb in (("native_bookkeep", False), ("torch_bookkeep", True)):
        w, eng = build(E, T, tb)
        c0 = wlib.launch_count()
        eng._rollout_eager()
        torch.cuda.synchronize()
        res[name] = {"libwdb200_launches_per_step": (wlib.launch_count() - c0) / T,
                     "ms_per_step": time_engine(eng, T)}
        res[name]["env_steps_per_s"] = E * 5 / (res[name]["ms_per_step"] * 1e-3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
