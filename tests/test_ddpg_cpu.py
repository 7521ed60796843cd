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
