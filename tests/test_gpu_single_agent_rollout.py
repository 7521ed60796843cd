"""GPU tests of the whole-rollout kernel for discrete single-agent envs
(wdb_single_agent_rollout, csrc/wdb_sa_rollout.cu; SURVEY.md section 8 rows a1-a11 for
BASELINE config 3).

The one-launch rollout must record exactly what the separate reference-shaped calls produce
when they are teacher-forced with its actions: the stand-alone step kernels (pinned bit for
bit to the reference's numba binaries in test_gpu_classic_control.py), the stand-alone
sampler (pinned in test_gpu_core.py), the generic bookkeeping and the reset kernel
(trainer_base.py:383-601 order: forward -> sample -> step -> bookkeep -> reset).  The in-kernel
float32 forward is checked against the torch module at 1e-5.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ENVS = {"cartpole": ("warp_drive_b200.envs.single_agent.cartpole", "CUDAClassicControlCartPoleEnv"),
        "mountain_car": ("warp_drive_b200.envs.single_agent.classic_control",
                         "CUDAClassicControlMountainCarEnv"),
        "acrobot": ("warp_drive_b200.envs.single_agent.classic_control",
                    "CUDAClassicControlAcrobotEnv")}


def _build(name, E, T, pool, fused, fc_dims=(32, 32), episode_length=30, seed=11):
    import importlib

    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.managers.function_manager import CUDASampler
    from warp_drive_b200.training.models.fully_connected import FullyConnected
    from warp_drive_b200.training.rollout import RolloutEngine
    from warp_drive_b200.training.utils.data_loader import create_and_push_data_placeholders

    mod, cls = ENVS[name]
    env = getattr(importlib.import_module(mod), cls)(
        episode_length=episode_length, env_backend="b200", reset_pool_size=pool, seed=seed)
    w = EnvWrapper(env, num_envs=E, env_backend="b200")
    w.reset_all_envs()
    pm = {"shared": [0]}
    s = CUDASampler(w.cuda_function_manager)
    create_and_push_data_placeholders(env_wrapper=w, action_sampler=s,
                                      policy_tag_to_agent_id_map=pm,
                                      training_batch_size_per_env=T)
    s.init_random(seed)
    if pool >= 2:
        w.init_reset_pool(seed=5)
    torch.manual_seed(0)
    cfg = {"type": "fully_connected", "fc_dims": list(fc_dims), "model_ckpt_filepath": ""}
    models = {"shared": FullyConnected(w, cfg, "shared", pm).cuda().eval()}
    # spread the initial states so that different replicas take different branches
    st = w.cuda_data_manager.data_on_device_via_torch("state")
    g = torch.Generator(device="cuda").manual_seed(3)
    st.add_(0.05 * (torch.rand(st.shape, device="cuda", generator=g) - 0.5))
    eng = RolloutEngine(w, models, pm, s, T, use_cuda_graph=False, use_fused_step=fused)
    return w, eng, s, models["shared"]


@pytest.mark.parametrize("name", ["cartpole", "mountain_car", "acrobot"])
@pytest.mark.parametrize("pool", [0, 8])
def test_one_launch_rollout_equals_separate_calls(name, pool):
    E, T = 192, 24
    wa, ea, sa, ma = _build(name, E, T, pool, fused=True)
    wb, eb, sb, mb = _build(name, E, T, pool, fused=False)
    assert ea.sa is not None, "the whole-rollout kernel is not in use"
    assert eb.sa is None and eb.fused is None
    dma, dmb = wa.cuda_data_manager, wb.cuda_data_manager
    A = ea.sa.A
    g = torch.Generator(device="cuda").manual_seed(9)
    n_done = 0
    for it in range(4):                      # 96 steps > 3 episodes of 30 steps
        u = torch.rand((T, E), device="cuda", generator=g).clamp_min(1e-7)
        probs = torch.zeros((T, E, A), device="cuda")
        ea.sa.launch(T, t0=0, record=True, uniforms=u, probs_out=probs)
        torch.cuda.synchronize()
        obs_b = dma.data_on_device_via_torch("processed_observations_batch_shared")
        act_b = dma.data_on_device_via_torch("sampled_actions_batch_shared")
        rew_b = dma.data_on_device_via_torch("rewards_batch_shared")
        done_b = dma.data_on_device_via_torch("done_flags_batch")
        for t in range(T):
            # the env B is in exactly the state A recorded
            assert torch.equal(dmb.data_on_device_via_torch("observations"), obs_b[t]), (it, t)
            with torch.no_grad():
                pr, _ = mb(dmb.data_on_device_via_torch("observations"))
            assert torch.allclose(pr[0].view(E, A), probs[t], atol=1e-5, rtol=1e-5), (it, t)
            # the stand-alone sampler on A's probabilities and uniforms draws A's actions
            sb.sample(dmb, probs[t].view(E, 1, A).contiguous(), "sampled_actions",
                      write_cum_distr=False, uniforms=u[t].contiguous())
            assert torch.equal(dmb.data_on_device_via_torch("sampled_actions"), act_b[t]), (it, t)
            wb.step_all_envs()
            eb.bookkeep(t)
            assert torch.equal(dmb.data_on_device_via_torch("rewards_batch_shared")[t], rew_b[t])
            assert torch.equal(dmb.data_on_device_via_torch("done_flags_batch")[t], done_b[t])
            n_done += int((done_b[t] > 0).sum())
            wb.reset_only_done_envs()
        for k in ("state", "observations", "_timestep_", "_done_", "rewards"):
            assert torch.equal(dma.data_on_device_via_torch(k), dmb.data_on_device_via_torch(k)), k
    assert n_done >= 2 * E
    assert int(ea.num_completed_episodes) == int(eb.num_completed_episodes) == n_done
    assert int(ea.episodic_step_sum) == int(eb.episodic_step_sum)
    ra, rb = float(ea.episodic_reward_sum["shared"]), float(eb.episodic_reward_sum["shared"])
    assert abs(ra - rb) <= 1e-3 * max(1.0, abs(rb))
    assert torch.equal(ea.step_running_sum, eb.step_running_sum)
    if name == "mountain_car":
        pass   # done == 2 (goal reached) is rare under a random policy; covered by the step tests


def test_device_rng_matches_the_sampler_stream():
    """Without injected uniforms the kernel draws from the sampler's Philox streams exactly
    like wdb_sample_actions (stream = env, one draw per timestep)."""
    E, T = 64, 8
    wa, ea, sa, ma = _build("cartpole", E, T, 0, fused=True)
    wb, eb, sb, mb = _build("cartpole", E, T, 0, fused=False)
    probs = torch.zeros((T, E, 2), device="cuda")
    ea.sa.launch(T, t0=0, record=True, probs_out=probs)
    act_b = wa.cuda_data_manager.data_on_device_via_torch("sampled_actions_batch_shared")
    dmb = wb.cuda_data_manager
    for t in range(T):
        sb.sample(dmb, probs[t].view(E, 1, 2).contiguous(), "sampled_actions",
                  write_cum_distr=False)
        assert torch.equal(dmb.data_on_device_via_torch("sampled_actions"), act_b[t]), t


def test_engine_uses_one_launch_per_rollout():
    from warp_drive_b200 import lib as wlib

    E, T = 256, 16
    w, e, s, m = _build("cartpole", E, T, 8, fused=True)
    e.rollout()
    torch.cuda.synchronize()
    c0 = wlib.launch_count()
    e.rollout()
    assert wlib.launch_count() - c0 == 1
    torch.cuda.synchronize()
    dm = w.cuda_data_manager
    acts = dm.data_on_device_via_torch("sampled_actions_batch_shared")
    assert int(acts.min()) >= 0 and int(acts.max()) <= 1
    assert float(dm.data_on_device_via_torch("rewards_batch_shared").min()) == 1.0


def test_wider_policy_and_three_hidden_layers():
    E, T = 64, 6
    wa, ea, sa, ma = _build("acrobot", E, T, 0, fused=True, fc_dims=(64, 48, 16))
    assert ea.sa is not None
    probs = torch.zeros((T, E, 3), device="cuda")
    ea.sa.launch(T, t0=0, record=True, probs_out=probs)
    obs_b = wa.cuda_data_manager.data_on_device_via_torch("processed_observations_batch_shared")
    with torch.no_grad():
        pr, _ = ma(obs_b)
    assert torch.allclose(pr[0].view(T, E, 3), probs, atol=1e-5, rtol=1e-5)
    # a policy too large for the per-thread forward falls back to the generic path
    wb, eb, sb, mb = _build("acrobot", E, T, 0, fused=True, fc_dims=(256, 256))
    assert eb.sa is None
    eb.rollout()
    torch.cuda.synchronize()
