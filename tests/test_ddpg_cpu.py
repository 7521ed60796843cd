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
