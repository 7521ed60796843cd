"""DDPG pieces (continuous-action training for Pendulum /
ContinuousMountainCar) on the CPU: the loss and the ring buffer are replayed against vectors
produced by the reference's own DDPG and RingBuffer classes (tests/golden/
make_ddpg_golden.py); the actor / critic modules are checked for the reference's parameter
names and forward contract."""
import numpy as np
import pytest
import torch

from conftest import load_golden


def test_ddpg_loss_matches_the_reference_class():
    from warp_drive_b200.training.algorithms.ddpg import DDPG

    g = load_golden("ddpg_reference.npz")
    for c in range(int(g["n_cases"])):
        T, E, Np, n, gamma, nr, na = g[f"case{c}__cfg"]
        algo = DDPG(discount_factor_gamma=float(gamma), normalize_advantage=bool(na),
                    normalize_return=bool(nr), n_step=int(n))
        t = {k: torch.from_numpy(g[f"case{c}__{k}"]) for k in (
            "rewards", "done", "values", "next_values", "j_values", "actions")}
        actor_loss, critic_loss, metrics = algo.compute_loss_and_metrics(
            100, t["actions"], t["rewards"], t["done"], t["values"], t["next_values"],
            t["j_values"], perform_logging=True)
        assert actor_loss.item() == pytest.approx(float(g[f"case{c}__actor_loss"]), rel=1e-5, abs=1e-6)
        assert critic_loss.item() == pytest.approx(float(g[f"case{c}__critic_loss"]), rel=1e-5, abs=1e-6)
        assert metrics["Mean (discounted) returns"] == pytest.approx(
            float(g[f"case{c}__mean_returns"]), rel=1e-5, abs=1e-6)
        assert metrics["Variance explained by the value function"] == pytest.approx(
            float(g[f"case{c}__var_explained"]), rel=1e-4, abs=1e-5)


def test_ring_buffer_matches_the_reference_class():
    from warp_drive_b200.training.utils.ring_buffer import RingBuffer, RingBufferManager

    g = load_golden("ddpg_reference.npz")
    items = torch.from_numpy(g["ring__items"])
    mgr = RingBufferManager()
    mgr.add("x", tensor=torch.zeros(5, 2, 3))
    assert mgr.has("x") and not mgr.has("y")
    ring = mgr.get("x")
    assert ring.unroll() is None and not ring.isfull()
    for i in range(items.shape[0]):
        ring.enqueue(items[i])
        assert np.array_equal(ring.unroll().numpy(), g[f"ring__unroll_{i}"]), i
        assert ring.isfull() == bool(g[f"ring__full_{i}"])
    with pytest.raises(AssertionError):
        RingBuffer(name="big", size=7, tensor=torch.zeros(6, 2))


class _StubEnvWrapper:
    """What the model classes read from an EnvWrapper (no device needed)."""

    class _DM:
        def get_shape(self, name=None):
            return (8, 4, 1, 3)

    def __init__(self, obs_dim=3, act_dim=1):
        from warp_drive_b200.utils import spaces

        class _Env:
            observation_space = {0: spaces.Box(-1.0, 1.0, shape=(obs_dim,), dtype=np.float32)}
            action_space = {0: spaces.Box(-2.0, 2.0, shape=(act_dim,), dtype=np.float32)}

        self.env = _Env()
        self.n_agents = 1
        self.cuda_data_manager = self._DM()


def test_actor_and_critic_follow_the_reference_contract():
    from warp_drive_b200.training.models.fully_connected import ModelFactory
    from warp_drive_b200.training.models.fully_connected_actor_critic import ActorAsPolicy

    env = _StubEnvWrapper()
    kw = dict(env=env, policy="shared", policy_tag_to_agent_id_map={"shared": [0]})
    actor = ModelFactory.create("fully_connected_actor")(
        model_config={"type": "fully_connected_actor", "fc_dims": [16, 8], "output_w": 2.0}, **kw)
    critic = ModelFactory.create("fully_connected_action_value_critic")(
        model_config={"type": "fully_connected_action_value_critic", "fc_dims": [16, 8]}, **kw)
    # parameter names of the reference checkpoints (fully_connected_actor_critic.py:31-41,99-109)
    assert sorted(actor.state_dict()) == [
        "fc.0.0.bias", "fc.0.0.weight", "fc.1.0.bias", "fc.1.0.weight",
        "policy_head.bias", "policy_head.weight"]
    assert sorted(critic.state_dict()) == [
        "fc.0.0.bias", "fc.0.0.weight", "fc.1.0.bias", "fc.1.0.weight",
        "vf_head.bias", "vf_head.weight"]
    assert critic.fc["0"][0].in_features == 3 + 1
    obs = torch.randn(5, 4, 1, 3)
    act = actor(obs)
    assert isinstance(act, list) and len(act) == 1 and act[0].shape == (5, 4, 1, 1)
    assert act[0].abs().max() <= 2.0                       # output_w * tanh
    with torch.no_grad():
        actor.policy_head.bias.fill_(50.0)
    assert torch.allclose(actor(obs)[0], torch.full((5, 4, 1, 1), 2.0))
    q = critic(obs, act)
    assert q.shape == (5, 4, 1)
    assert torch.equal(q, critic(obs, act[0]))             # list or tensor action
    q.sum().backward()                                       # dJ/d(actor) flows through Q
    assert actor.policy_head.weight.grad is not None
    probs, values = ActorAsPolicy(actor)(obs)
    assert values is None and torch.equal(probs[0], actor(obs)[0])


def test_a2c_negative_positive_env_sampling():
    """single_mountain_car.yaml's `neg_pos_env_ratio` (a2c.py:58-69,196-220 of the reference):
    envs that reached the goal (done == 2) are all kept, the others down-sampled."""
    from warp_drive_b200.training.algorithms.policygradient import A2C

    torch.manual_seed(0)
    T, E, Np = 6, 40, 1
    done = torch.zeros(T, E, dtype=torch.int32)
    done[3, [2, 17, 30]] = 2          # three envs reached the goal
    done[5, 8] = 1                    # an episode end is not a positive
    keep, n_pos, n_neg = A2C._sample_positive_negative_env_ids(done, 5)
    assert n_pos == 3 and n_neg == 15 and keep.numel() == 18
    assert set(keep[:3].tolist()) == {2, 17, 30} and len(set(keep.tolist())) == 18
    assert A2C._sample_positive_negative_env_ids(done, 50)[0] is None     # nothing to drop
    assert A2C._sample_positive_negative_env_ids(torch.zeros_like(done), 5)[0] is None
    algo = A2C(discount_factor_gamma=0.99, vf_loss_coeff=1.0, entropy_coeff=0.05)
    probs = [torch.softmax(torch.randn(T, E, Np, 3), -1).requires_grad_()]
    values = torch.randn(T, E, Np, requires_grad=True)
    loss, metrics = algo.compute_loss_and_metrics(
        timestep=10, actions_batch=torch.randint(0, 3, (T, E, Np, 1)),
        rewards_batch=-torch.ones(T, E, Np), done_flags_batch=done,
        action_probabilities_batch=probs, value_functions_batch=values,
        perform_logging=True, negative_positive_ratio=5)
    assert metrics["Num of Positive Sampled Envs"] == 3
    assert metrics["Num of Negative Sampled Envs"] == 15
    loss.backward()
    touched = (values.grad.abs().sum(dim=(0, 2)) > 0).sum().item()
    assert touched == 18               # only the sampled envs contribute
