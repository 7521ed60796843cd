#!/usr/bin/env python
"""Learning-curve probe behind tests/test_gpu_update.py::test_cartpole_learns_*: mean episodic
steps of CartPole per training iteration for several seeds, with and without TF32 matmuls."""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import yaml  # noqa: E402


def run(seed, tf32, iters=150):
    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.single_agent.cartpole import CUDAClassicControlCartPoleEnv
    from warp_drive_b200.training.trainer import Trainer

    torch.backends.cuda.matmul.allow_tf32 = tf32
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "warp_drive_b200", "training", "run_configs",
                           "single_cartpole.yaml"), encoding="utf8") as fp:
        cfg = yaml.safe_load(fp)
    E, T = 256, 64
    cfg["env"].update(episode_length=200, reset_pool_size=64)
    cfg["policy"]["shared"].update(lr=0.01, entropy_coeff=0.01)
    cfg["trainer"].update(num_envs=E, train_batch_size=E * T, num_episodes=10 ** 6, seed=seed)
    cfg["saving"].update(basedir=tempfile.mkdtemp(), metrics_log_freq=10 ** 9,
                         model_params_save_freq=10 ** 9)
    env = CUDAClassicControlCartPoleEnv(**cfg["env"])
    w = EnvWrapper(env, num_envs=E, env_backend="b200")
    tr = Trainer(w, cfg, {"shared": [0]}, verbose=False)
    w.reset_all_envs()
    tr.engine.resync_observations()
    means = []
    for it in range(iters):
        tr._generate_rollout_batch()
        tr._update_model_params(it)
        n = int(tr.engine.num_completed_episodes)
        means.append(float(tr.engine.episodic_step_sum) / max(n, 1))
        tr.engine.episodic_step_sum.zero_()
        tr.engine.num_completed_episodes.zero_()
        tr.engine.episodic_reward_sum["shared"].zero_()
    return {"seed": seed, "tf32": tf32, "first": float(np.mean(means[1:6])),
            "last": float(np.mean(means[-5:])), "curve": [round(m, 1) for m in means[::10]]}


if __name__ == "__main__":
    for tf32 in (False, True):
        for seed in (3, 4, 5, 6):
            print(json.dumps(run(seed, tf32)), flush=True)
