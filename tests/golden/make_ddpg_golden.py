"""Generate tests/golden/ddpg_reference.npz with the REFERENCE's own classes.

Run in the build container only (needs /root/reference; CPU torch is enough):

    python tests/golden/make_ddpg_golden.py

Imports warp_drive.training.algorithms.policygradient.ddpg.DDPG and
warp_drive.training.utils.ring_buffer.RingBuffer from /root/reference (nothing is copied),
feeds them seeded random batches / enqueue sequences and records inputs and outputs.
tests/test_ddpg_cpu.py replays the same inputs through warp_drive_b200's classes.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

from warp_drive.training.algorithms.policygradient.ddpg import DDPG  # noqa: E402
from warp_drive.training.utils.ring_buffer import RingBuffer  # noqa: E402

CASES = [  # (T, E, Np, n_step, gamma, normalize_return, normalize_advantage)
    (12, 5, 3, 1, 0.99, False, False),
    (12, 5, 3, 3, 0.99, False, False),
    (9, 4, 1, 5, 0.9, True, False),
    (7, 6, 2, 7, 1.0, False, True),
    (10, 3, 2, 2, 0.5, True, True),
]


class _StubDataManager:
    def __init__(self, tensor):
        self.tensor = tensor

    def is_data_on_device_via_torch(self, name):
        return True

    def data_on_device_via_torch(self, name=None):
        return self.tensor

    def get_shape(self, name):
        return tuple(self.tensor.shape)


def main():
    out = {}
    g = torch.Generator().manual_seed(1234)
    for c, (T, E, Np, n, gamma, nr, na) in enumerate(CASES):
        rewards = torch.randn(T, E, Np, generator=g)
        done = (torch.rand(T, E, generator=g) < 0.2).to(torch.int32)
        values = torch.randn(T, E, Np, generator=g)
        next_values = torch.randn(T - 1, E, Np, generator=g)
        j_values = torch.randn(T, E, Np, generator=g)
        actions = torch.randn(T, E, Np, 1, generator=g)
        algo = DDPG(discount_factor_gamma=gamma, normalize_advantage=na, normalize_return=nr,
                    n_step=n)
        actor_loss, critic_loss, metrics = algo.compute_loss_and_metrics(
            100, actions, rewards, done, values, next_values, j_values, perform_logging=True)
        for k, v in dict(rewards=rewards, done=done, values=values, next_values=next_values,
                         j_values=j_values, actions=actions).items():
            out[f"case{c}__{k}"] = v.numpy()
        out[f"case{c}__cfg"] = np.array([T, E, Np, n, gamma, nr, na], np.float64)
        out[f"case{c}__actor_loss"] = np.float64(actor_loss.item())
        out[f"case{c}__critic_loss"] = np.float64(critic_loss.item())
        out[f"case{c}__mean_returns"] = np.float64(metrics["Mean (discounted) returns"])
        out[f"case{c}__var_explained"] = np.float64(
            metrics["Variance explained by the value function"])
    # ring buffer managing its whole 5-slot container (how the trainer uses it,
    # trainer_ddpg.py:91-94), 13 enqueues, unroll after each
    container = torch.zeros(5, 2, 3)
    ring = RingBuffer(name="x", data_manager=_StubDataManager(container))
    items = torch.randn(13, 2, 3, generator=g)
    out["ring__items"] = items.numpy()
    for i in range(13):
        ring.enqueue(items[i])
        out[f"ring__unroll_{i}"] = ring.unroll().clone().numpy()
        out[f"ring__full_{i}"] = np.bool_(ring.isfull())
    out["n_cases"] = np.int32(len(CASES))
    path = os.path.join(HERE, "ddpg_reference.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
