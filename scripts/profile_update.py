#!/usr/bin/env python
"""Kernel-level breakdown of ONE A2C/PPO update at BASELINE config 2 (T = 10) with
torch.profiler: top CUDA kernels by total time.  Usage: python scripts/profile_update.py"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    args = argparse.Namespace(envs=2000, train_steps=10, algo="PPO")
    import copy

    import yaml

    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.tag_continuous import TagContinuous
    from warp_drive_b200.training.trainer import Trainer

    torch.backends.cuda.matmul.allow_tf32 = True
    E, T = args.envs, args.train_steps
    with open(os.path.join(bench.ROOT, "warp_drive_b200", "training", "run_configs",
                           "tag_continuous.yaml"), encoding="utf8") as fp:
        cfg = yaml.safe_load(fp)
    cfg["env"].update(bench.ENV_CONFIG)
    cfg["trainer"].update(num_envs=E, train_batch_size=E * T, num_episodes=10 ** 6, seed=1234)
    for p in cfg["policy"].values():
        p["algorithm"] = args.algo
        p["clip_param"] = 0.1
    cfg["saving"].update(metrics_log_freq=10 ** 9, model_params_save_freq=10 ** 9,
                         basedir="/tmp", name="prof_update", tag="x")
    env = TagContinuous(**cfg["env"])
    wrapper = EnvWrapper(env, num_envs=E, env_backend="b200")
    pm = {"runner": sorted(env.runners), "tagger": sorted(env.taggers)}
    trainer = Trainer(wrapper, copy.deepcopy(cfg), pm, verbose=False)
    wrapper.reset_all_envs()
    trainer.engine.resync_observations()
    for i in range(3):
        trainer._generate_rollout_batch()
        trainer._update_model_params(1000 + i)
    torch.cuda.synchronize()
    trainer._generate_rollout_batch()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        trainer._update_model_params(2000)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if e.device_type.name == "CUDA" or t:
            rows.append((t, e.count, e.key))
    rows = [r for r in rows if r[0] > 0]
    rows.sort(reverse=True)
    seen = set()
    print("total_us count name")
    for t, c, k in rows[:45]:
        if k in seen:
            continue
        seen.add(k)
        print(f"{t:10.1f} {c:5d} {k[:150]}")


if __name__ == "__main__":
    main()
