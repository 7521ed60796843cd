"""GPU tests of the fused update path: the one-kernel A2C / PPO
loss + gradients (csrc/wdb_update.cu, reference a2c.py:80-130 / ppo.py:82-141) against the
float32 torch expression of the same loss at 1e-5, the flat-arena Adam + clipping against
torch.optim.Adam + clip_grad_norm_, and a learning assertion on CartPole."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(T, E, Np, heads, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    logits = [torch.randn((T, E, Np, a), device="cuda", generator=g).mul_(1.5).requires_grad_()
              for a in heads]
    values = torch.randn((T, E, Np), device="cuda", generator=g).requires_grad_()
    actions = torch.stack([torch.randint(0, a, (T, E, Np), device="cuda", generator=g)
                           for a in heads], -1)
    rewards = torch.randn((T, E, Np), device="cuda", generator=g)
    done = (torch.rand((T, E), device="cuda", generator=g) < 0.15).to(torch.int32)
    done[2, 0] = 2            # MountainCar's "goal reached" flag counts as done
    return logits, values, actions, rewards, done


@pytest.mark.parametrize("algo", ["A2C", "PPO"])
@pytest.mark.parametrize("heads", [(21, 21), (2,), (5, 3, 4)])
def test_fused_loss_matches_torch_expression(algo, heads):
    from warp_drive_b200.training.algorithms.policygradient import A2C, PPO

    T, E, Np = 7, 6, 9
    kw = dict(discount_factor_gamma=0.97, vf_loss_coeff=0.7, entropy_coeff=0.03)
    make = (lambda: A2C(**kw)) if algo == "A2C" else (lambda: PPO(clip_param=0.1, **kw))
    out = {}
    for fused in (False, True):
        logits, values, actions, rewards, done = _batch(T, E, Np, heads, seed=5)
        probs = [torch.softmax(lg, -1) for lg in logits]
        tr = make()
        tr.use_fused_loss = fused
        loss, metrics = tr.compute_loss_and_metrics(
            100, actions.long(), rewards, done, probs, values, perform_logging=True)
        loss.backward()
        out[fused] = (loss.detach(), [lg.grad.clone() for lg in logits], values.grad.clone(),
                      metrics)
    (l0, g0, v0, m0), (l1, g1, v1, m1) = out[False], out[True]
    assert torch.allclose(l0, l1, atol=1e-5, rtol=1e-5), (float(l0), float(l1))
    for a, b in zip(g0, g1):
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-4), float((a - b).abs().max())
    assert torch.allclose(v0, v1, atol=1e-6, rtol=1e-4)
    for k in ("Policy loss", "Value function loss", "Mean entropy", "Mean (discounted) returns",
              "Mean advantages", "Variance explained by the value function"):
        assert abs(m0[k] - m1[k]) <= 1e-5 + 1e-4 * abs(m0[k]), (k, m0[k], m1[k])


def test_fused_loss_handles_clamped_probabilities():
    """Probabilities below torch's clamp (eps) get zero log-prob gradient, as in
    torch.distributions.Categorical."""
    from warp_drive_b200.training.algorithms.policygradient import A2C

    T, E, Np = 3, 2, 4
    res = {}
    for fused in (False, True):
        g = torch.Generator(device="cuda").manual_seed(1)
        p = torch.rand((T, E, Np, 4), device="cuda", generator=g)
        p[..., 0] = 0.0                       # an impossible action ...
        p = (p / p.sum(-1, keepdim=True)).requires_grad_()
        values = torch.zeros((T, E, Np), device="cuda", requires_grad=True)
        actions = torch.randint(1, 4, (T, E, Np, 1), device="cuda", generator=g)
        rewards = torch.ones((T, E, Np), device="cuda")
        done = torch.zeros((T, E), dtype=torch.int32, device="cuda")
        tr = A2C(discount_factor_gamma=0.9, vf_loss_coeff=0.5, entropy_coeff=0.01)
        tr.use_fused_loss = fused
        loss, _ = tr.compute_loss_and_metrics(1, actions.long(), rewards, done, [p], values)
        loss.backward()
        res[fused] = (loss.detach(), p.grad.clone())
    assert torch.allclose(res[False][0], res[True][0], atol=1e-5)
    # compare after projecting out the per-row constant that a softmax backward removes
    a, b = res[False][1], res[True][1]
    assert torch.isfinite(b).all()
    assert torch.allclose(a, b, atol=1e-5, rtol=1e-4), float((a - b).abs().max())


def test_flat_adam_matches_torch_adam_with_clipping():
    from warp_drive_b200.training.utils.flat_adam import FlatAdam

    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(13, 32), torch.nn.ReLU(), torch.nn.Linear(32, 5)).cuda()
    mine = copy.deepcopy(ref)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=3e-3)
    opt = FlatAdam(mine.parameters(), lr=3e-3)
    g = torch.Generator(device="cuda").manual_seed(2)
    for step in range(8):
        x = torch.randn((64, 13), device="cuda", generator=g)
        y = torch.randn((64, 5), device="cuda", generator=g)
        for model, o in ((ref, opt_ref), (mine, opt)):
            o.zero_grad()
            (model(x) - y).square().mean().mul(50.0).backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        opt_ref.step()
        opt.step(max_grad_norm=0.5)
        for a, b in zip(ref.parameters(), mine.parameters()):
            assert torch.allclose(a, b, atol=2e-6, rtol=1e-5), (step, float((a - b).abs().max()))
    # without clipping, and the lr handle
    opt.param_groups[0]["lr"] = 1e-3
    for gparam in opt_ref.param_groups:
        gparam["lr"] = 1e-3
    for model, o in ((ref, opt_ref), (mine, opt)):
        o.zero_grad()
        model(x).square().mean().backward()
        o.step()
    for a, b in zip(ref.parameters(), mine.parameters()):
        assert torch.allclose(a, b, atol=2e-6, rtol=1e-5)
    # parameters are views of one arena
    assert all(p.data_ptr() >= opt.params.data_ptr() for p in mine.parameters())


def test_cartpole_learns_with_the_fused_rollout_and_update(tmp_path):
    """Learning assertion: 'Mean episodic steps' of CartPole rises while training with the
    whole-rollout kernel + fused loss + flat Adam (a random policy balances ~22 steps)."""
    import yaml

    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.single_agent.cartpole import CUDAClassicControlCartPoleEnv
    from warp_drive_b200.training.trainer import Trainer

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "warp_drive_b200", "training", "run_configs",
                           "single_cartpole.yaml"), encoding="utf8") as fp:
        cfg = yaml.safe_load(fp)
    E, T, iters = 256, 64, 150
    cfg["env"].update(episode_length=200, reset_pool_size=64)
    cfg["policy"]["shared"].update(lr=0.01, entropy_coeff=0.01)
    cfg["trainer"].update(num_envs=E, train_batch_size=E * T, num_episodes=10 ** 6, seed=3)
    cfg["saving"].update(basedir=str(tmp_path), metrics_log_freq=10 ** 9,
                         model_params_save_freq=10 ** 9)
    env = CUDAClassicControlCartPoleEnv(**cfg["env"])
    w = EnvWrapper(env, num_envs=E, env_backend="b200")
    tr = Trainer(w, cfg, {"shared": [0]}, verbose=False)
    assert tr.engine.sa is not None
    w.reset_all_envs()
    tr.engine.resync_observations()
    means = []
    tf32_before = torch.backends.cuda.matmul.allow_tf32     # other tests may have switched it on
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for it in range(iters):
            tr._generate_rollout_batch()
            tr._update_model_params(it)
            n = int(tr.engine.num_completed_episodes)
            means.append(float(tr.engine.episodic_step_sum) / max(n, 1))
            tr.engine.episodic_step_sum.zero_()
            tr.engine.num_completed_episodes.zero_()
            tr.engine.episodic_reward_sum["shared"].zero_()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32_before
    # (iteration 0 is a logging iteration: the trainer itself clears the counters there)
    # The curve is noisy at this learning rate (scripts/cartpole_learning_probe.py: 8 runs, seeds
    # 3-6 with and without TF32, every one reaches >= 45 steps for stretches but may dip back):
    # compare the best 10-iteration stretch with the random-policy level of the first iterations.
    first = np.mean(means[1:6])
    best = max(np.mean(means[i:i + 10]) for i in range(10, iters - 9))
    assert best > 1.5 * first, (first, best, [round(m, 1) for m in means[::10]])


@pytest.mark.parametrize("heads", [(21, 21), (3,), (40, 30)])
@pytest.mark.parametrize("rows_shape", [(7, 13, 5), (1000,)])
def test_fused_train_forward_matches_module_autograd(heads, rows_shape):
    """models/fused_mlp_train.py (bias + ReLU in the GEMM epilogue, one GEMM for all heads,
    fused softmax / ReLU-backward / bias-gradient kernels) against the plain torch module under
    autograd: same outputs and the same parameter gradients (float32, sums in another order)."""
    from test_gpu_mlp import _Model

    F, H = 71, 64
    torch.manual_seed(5)
    m = _Model(F, H, heads[0], heads[1] if len(heads) > 1 else 1).cuda()
    if len(heads) == 1:
        m.policy_head = torch.nn.ModuleList([m.policy_head[0]])
        m.output_dims = [heads[0]]
    m.action_mask = None
    from warp_drive_b200.training.models import fused_mlp_train

    obs = torch.randn(*rows_shape, F, device="cuda")
    assert fused_mlp_train.supported(m, obs)
    g = torch.Generator(device="cuda").manual_seed(1)

    def loss_of(probs, values):
        out = (values * wv).sum()
        for p, w in zip(probs, wp):
            out = out + (p * w).sum()
        return out

    probs_f, values_f = fused_mlp_train.fused_train_forward(m, obs)
    wp = [torch.randn(p.shape, device="cuda", generator=g) for p in probs_f]
    wv = torch.randn(values_f.shape, device="cuda", generator=g)
    loss_of(probs_f, values_f).backward()
    grads_f = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad()
    x = m.fc["1"](m.fc["0"](obs))
    probs_t = [torch.softmax(h(x), -1) for h in m.policy_head]
    values_t = m.vf_head(x)[..., 0]
    loss_of(probs_t, values_t).backward()
    for a, b in zip(probs_f, probs_t):
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-6, rtol=1e-5)
    assert values_f.shape == values_t.shape and torch.allclose(values_f, values_t, atol=1e-5, rtol=1e-5)
    for n, p in m.named_parameters():
        assert torch.allclose(grads_f[n], p.grad, atol=2e-4, rtol=1e-4), (
            n, float((grads_f[n] - p.grad).abs().max()))


def test_trainer_update_uses_the_fused_train_forward(tmp_path):
    """FullyConnected.forward under autograd goes through the fused node (and stays on the plain
    module path under no_grad / for models it does not cover)."""
    from test_gpu_training import _run_config
    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.envs.tag_continuous import TagContinuous
    from warp_drive_b200.training.trainer import Trainer

    cfg = _run_config("tag_continuous", num_envs=8, train_batch_size=8 * 4, num_episodes=4)
    cfg["saving"]["basedir"] = str(tmp_path)
    env = TagContinuous(**cfg["env"])
    w = EnvWrapper(env, num_envs=8, env_backend="b200")
    tr = Trainer(w, cfg, {"runner": sorted(env.runners), "tagger": sorted(env.taggers)},
                 verbose=False)
    model = tr.models["runner"]
    obs = torch.randn(4, 8, len(env.runners), model.flattened_obs_size, device="cuda")
    probs, values = model(obs)
    node = probs[0].grad_fn                    # the [T, E, Np, A] view of the fused node's output
    assert node is not None and "FusedMLPTrain" in type(node.next_functions[0][0]).__name__
    with torch.no_grad():
        probs_ng, values_ng = model(obs)
    assert torch.allclose(probs[0], probs_ng[0], atol=1e-6) and torch.allclose(values, values_ng, atol=1e-5)
    tr.graceful_close()
