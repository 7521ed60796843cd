"""GPU end-to-end training tests: the reference's own criterion is "runs without an exception"
(tests/wd_training/pycuda_tests/test_env_training.py:56-76, tag_gridworld + a shrunken
tag_continuous); here additionally: losses are finite, parameters move, the rollout batch
is consistent, checkpoints round-trip, and the returns kernel equals the host recursion."""
import copy
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_config(name, **trainer_overrides):
    from warp_drive_b200.training.trainer import load_run_config

    cfg = load_run_config(name)
    cfg["trainer"].update(trainer_overrides)
    cfg["saving"].update(metrics_log_freq=1, model_params_save_freq=2)
    return cfg


def _snapshot(models):
    return {p: [q.detach().clone() for q in m.parameters()] for p, m in models.items()}


def _train_and_check(trainer):
    before = _snapshot(trainer.models)
    trainer.train()
    for p in trainer.policies_to_train:
        moved = any(not torch.equal(a, b) for a, b in
                    zip(before[p], trainer.models[p].parameters()))
        assert moved, f"policy {p} did not train"
        for q in trainer.models[p].parameters():
            assert torch.isfinite(q).all()
    results = os.path.join(trainer.save_dir, "results.json")
    assert os.path.exists(results)
    import json

    lines = [json.loads(l) for l in open(results)]
    assert len(lines) == trainer.num_iters
    for rec in lines:
        for p in trainer.policies_to_train:
            assert np.isfinite(rec[p]["Total loss"]) and np.isfinite(rec[p]["Mean entropy"])
    ckpts = glob.glob(os.path.join(trainer.save_dir, "*.state_dict"))
    assert len(ckpts) >= len(trainer.policies)
    return lines


def test_train_tag_continuous_fused_rollouts(tmp_path):
    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.tag_continuous import TagContinuous
    from warp_drive_b200.training.trainer import Trainer

    cfg = _run_config("tag_continuous", num_envs=32, train_batch_size=32 * 25, num_episodes=96)
    cfg["env"].update(num_taggers=2, num_runners=10, episode_length=50)
    cfg["saving"]["basedir"] = str(tmp_path)
    env = TagContinuous(**cfg["env"])
    wrapper = EnvWrapper(env, num_envs=32, env_backend="pycuda")
    pm = {"runner": sorted(env.runners), "tagger": sorted(env.taggers)}
    trainer = Trainer(env_wrapper=wrapper, config=cfg, policy_tag_to_agent_id_map=pm,
                      results_dir="t", verbose=False)
    assert trainer.engine.fused is not None            # the single-launch timestep is in use
    lines = _train_and_check(trainer)
    assert lines[-1]["runner"]["Mean episodic steps"] <= 50
    # batch consistency: rewards of the batch are what the kernel reported, actions in range
    dm = wrapper.cuda_data_manager
    acts = dm.data_on_device_via_torch("sampled_actions_batch_runner")
    assert int(acts.min()) >= 0 and int(acts.max()) <= 20
    done = dm.data_on_device_via_torch("done_flags_batch")
    assert set(done.unique().tolist()) <= {0, 1}
    # checkpoint round trip through the reference's file naming
    step_of = lambda f: int(os.path.basename(f).split(".state_dict")[0].split("_")[-1])  # noqa: E731
    ck = max(glob.glob(os.path.join(trainer.save_dir, "runner_*.state_dict")), key=step_of)
    cfg2 = copy.deepcopy(cfg)
    cfg2["policy"]["runner"]["model"]["model_ckpt_filepath"] = ck
    env2 = TagContinuous(**cfg["env"])
    w2 = EnvWrapper(env2, num_envs=32, env_backend="numba")
    t2 = Trainer(env_wrapper=w2, config=cfg2, policy_tag_to_agent_id_map=pm,
                 results_dir="t2", verbose=False)
    assert t2.current_timestep["runner"] == step_of(ck) == trainer.current_timestep["runner"]
    for a, b in zip(trainer.models["runner"].parameters(), t2.models["runner"].parameters()):
        assert torch.equal(a.cpu(), b.cpu())
    states = t2.fetch_episode_states(["loc_x", "loc_y", "still_in_the_game"], env_id=1,
                                     include_rewards_actions=True)
    assert states["loc_x"].shape[1] == 12 and states["loc_x"].shape[0] >= 2
    assert np.isfinite(states["loc_x"]).all()
    # the device-side episode log (one pull at the end) returns exactly what the reference-
    # shaped per-step host pulls return
    names = ["loc_x", "loc_y", "speed", "still_in_the_game", "observations"]
    t2.cuda_sample_controller.init_random(77)
    dev = t2.fetch_episode_states(names, env_id=3, include_rewards_actions=True,
                                  include_probabilities=True)
    t2.cuda_sample_controller.init_random(77)
    host = t2._fetch_episode_states_host_pull(names, env_id=3, include_rewards_actions=True,
                                              include_probabilities=True)
    assert set(dev) == set(host)
    for k in names + ["sampled_actions", "rewards"]:
        assert dev[k].shape == host[k].shape, (k, dev[k].shape, host[k].shape)
        assert np.array_equal(dev[k], host[k], equal_nan=True), k
    assert len(dev["probabilities"]) == len(host["probabilities"])
    for pa, pb in zip(dev["probabilities"], host["probabilities"]):
        for x, y in zip(pa, pb):
            assert np.array_equal(x, y)
    trainer.graceful_close()


def test_train_tag_gridworld_generic_path(tmp_path):
    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.tag_gridworld import CUDATagGridWorld
    from warp_drive_b200.training.trainer import Trainer

    cfg = _run_config("tag_gridworld", num_envs=50, train_batch_size=50 * 20, num_episodes=100)
    cfg["env"].update(grid_length=10, episode_length=40)
    cfg["saving"]["basedir"] = str(tmp_path)
    env = CUDATagGridWorld(**cfg["env"])
    wrapper = EnvWrapper(env, num_envs=50, env_backend="pycuda")
    trainer = Trainer(env_wrapper=wrapper, config=cfg,
                      policy_tag_to_agent_id_map={"shared": list(range(env.num_agents))},
                      results_dir="g", verbose=False)
    assert trainer.engine.fused is None
    _train_and_check(trainer)


def test_train_cartpole_with_reset_pool(tmp_path):
    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.single_agent.cartpole import CUDAClassicControlCartPoleEnv
    from warp_drive_b200.training.trainer import Trainer

    cfg = _run_config("single_cartpole", num_envs=256, train_batch_size=256 * 16,
                      num_episodes=164)
    cfg["env"].update(episode_length=100, reset_pool_size=64)
    cfg["saving"]["basedir"] = str(tmp_path)
    env = CUDAClassicControlCartPoleEnv(**cfg["env"])
    wrapper = EnvWrapper(env, num_envs=256, env_backend="numba")
    trainer = Trainer(env_wrapper=wrapper, config=cfg,
                      policy_tag_to_agent_id_map={"shared": [0]}, results_dir="c",
                      verbose=False)
    lines = _train_and_check(trainer)
    # a random-ish policy drops the pole well before 100 steps
    assert 5 < lines[-1]["shared"]["Mean episodic steps"] < 100


def test_discounted_returns_kernel_matches_recursion():
    from warp_drive_b200.training.algorithms.policygradient import discounted_returns

    g = torch.Generator().manual_seed(0)
    T, E, Np = 37, 19, 11
    rewards = torch.randn(T, E, Np, generator=g)
    values = torch.randn(T, E, Np, generator=g)
    done = (torch.rand(T, E, generator=g) < 0.1).int()
    want = discounted_returns(rewards, done, values, 0.98)          # host recursion
    got = discounted_returns(rewards.cuda(), done.cuda(), values.cuda(), 0.98).cpu()
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


# --------------------------------------------------------------- classic control (8 f2)
@pytest.mark.parametrize("config,cls", [
    ("single_acrobot", "CUDAClassicControlAcrobotEnv"),
    ("single_mountain_car", "CUDAClassicControlMountainCarEnv")])
def test_train_discrete_classic_control_with_the_reference_configs(config, cls, tmp_path):
    """single_acrobot.yaml / single_mountain_car.yaml (values of the reference's files) through
    the A2C trainer; MountainCar additionally exercises `neg_pos_env_ratio` (done == 2)."""
    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.single_agent import classic_control as cc
    from warp_drive_b200.training.trainer import Trainer

    cfg = _run_config(config, num_envs=128, train_batch_size=128 * 20, num_episodes=256)
    cfg["env"].update(episode_length=40, reset_pool_size=32)
    cfg["saving"]["basedir"] = str(tmp_path)
    env = getattr(cc, cls)(**cfg["env"])
    wrapper = EnvWrapper(env, num_envs=128, env_backend="numba")
    trainer = Trainer(env_wrapper=wrapper, config=cfg,
                      policy_tag_to_agent_id_map={"shared": [0]}, results_dir="x", verbose=False)
    lines = _train_and_check(trainer)
    assert lines[-1]["shared"]["Mean episodic steps"] <= 40
    if config == "single_mountain_car":
        assert "Num of Positive Sampled Envs" in lines[-1]["shared"]


@pytest.mark.parametrize("config,cls", [
    ("single_pendulum", "CUDAClassicControlPendulumEnv"),
    ("single_continuous_mountain_car", "CUDAClassicControlContinuousMountainCarEnv")])
def test_train_ddpg_on_continuous_classic_control(config, cls, tmp_path):
    """single_pendulum.yaml / single_continuous_mountain_car.yaml through TrainerDDPG: actor
    forward -> OU sampler kernel -> env step kernel -> replay ring -> n-step DDPG update."""
    import json

    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.single_agent import classic_control as cc
    from warp_drive_b200.training.trainer_ddpg import TrainerDDPG

    E, T = 256, 8
    cfg = _run_config(config, num_envs=E, train_batch_size=E * T, num_episodes=E * 2)
    cfg["env"].update(episode_length=32, reset_pool_size=64)
    cfg["saving"]["basedir"] = str(tmp_path)
    env = getattr(cc, cls)(**cfg["env"])
    wrapper = EnvWrapper(env, num_envs=E, env_backend="numba")
    trainer = TrainerDDPG(env_wrapper=wrapper, config=cfg,
                          policy_tag_to_agent_id_map={"shared": [0]}, results_dir="d",
                          verbose=False)
    assert trainer.n_step == 5 and trainer.engine.ring_capacity == T + 4
    assert trainer.num_iters == 8
    actor0 = [p.detach().clone() for p in trainer.actor_models["shared"].parameters()]
    critic0 = [p.detach().clone() for p in trainer.critic_models["shared"].parameters()]
    target0 = [p.detach().clone() for p in trainer.target_actor_models["shared"].parameters()]
    trainer.train()
    for before, model in ((actor0, trainer.actor_models), (critic0, trainer.critic_models),
                          (target0, trainer.target_actor_models)):
        after = list(model["shared"].parameters())
        assert any(not torch.equal(a, b) for a, b in zip(before, after))
        assert all(torch.isfinite(b).all() for b in after)
    # the target net trails the online net (tau = 0.05), it is not a copy of it
    assert any(not torch.equal(a, b) for a, b in zip(
        trainer.actor_models["shared"].parameters(),
        trainer.target_actor_models["shared"].parameters()))
    lines = [json.loads(l) for l in open(os.path.join(trainer.save_dir, "results.json"))]
    assert len(lines) == trainer.num_iters
    trained = [rec["shared"] for rec in lines if "Critic loss" in rec["shared"]]
    assert len(trained) == trainer.num_iters - 1       # iteration 0 only fills the ring
    for rec in trained:
        assert np.isfinite(rec["Actor loss"]) and np.isfinite(rec["Critic loss"])
    # actions carry OU noise around the actor output and stay finite
    dm = wrapper.cuda_data_manager
    acts = dm.data_on_device_via_torch("sampled_actions_batch_shared")
    assert acts.dtype == torch.float32 and torch.isfinite(acts).all() and acts.std() > 0
    # the ring holds a time-ordered window: timesteps of one env advance by one (mod resets)
    ckpts = sorted(os.path.basename(p) for p in glob.glob(os.path.join(trainer.save_dir, "*.state_dict")))
    assert any(c.startswith("shared_actor_") for c in ckpts)
    assert any(c.startswith("shared_critic_") for c in ckpts)
    if config == "single_continuous_mountain_car":      # evaluator: True in that config
        assert "Mean episodic reward (test)" in lines[-1]["shared"]
    # checkpoints round-trip through load_model_checkpoint
    ts = trainer.current_timestep["shared"]
    paths = {"shared": {k: os.path.join(trainer.save_dir, f"shared_{k}_{ts}.state_dict")
                        for k in ("actor", "critic")}}
    with torch.no_grad():
        for p in trainer.actor_models["shared"].parameters():
            p.zero_()
    trainer.load_model_checkpoint(paths)
    assert any(p.abs().sum() > 0 for p in trainer.actor_models["shared"].parameters())


def test_evaluate_episodes_counts_one_episode_per_env(tmp_path):
    """Trainer.evaluate_episodes (trainer_base.py:794-846 of the reference): one episode per
    env replica, reward / step sums accumulated on the device.  CartPole pays 1 per step, so
    the two sums must agree env by env."""
    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.single_agent.cartpole import CUDAClassicControlCartPoleEnv
    from warp_drive_b200.training.trainer import Trainer

    cfg = _run_config("single_cartpole", num_envs=128, train_batch_size=128 * 10,
                      num_episodes=64)
    cfg["env"].update(episode_length=60, reset_pool_size=16)
    cfg["saving"]["basedir"] = str(tmp_path)
    env = CUDAClassicControlCartPoleEnv(**cfg["env"])
    wrapper = EnvWrapper(env, num_envs=128, env_backend="numba")
    trainer = Trainer(env_wrapper=wrapper, config=cfg,
                      policy_tag_to_agent_id_map={"shared": [0]}, results_dir="e",
                      verbose=False)
    reward_sum, step_sum = trainer.evaluate_episodes()
    steps = step_sum["shared"].cpu().numpy()
    rewards = reward_sum["shared"].cpu().numpy()[:, 0]
    assert ((steps >= 1) & (steps <= 60)).all()
    assert np.array_equal(rewards, steps.astype(np.float32))
    assert 5 < steps.mean() < 60          # an untrained policy drops the pole early
    # the training state is usable afterwards
    dm = wrapper.cuda_data_manager
    assert not dm.data_on_device_via_torch("_done_").any()
    trainer.train()
