#!/usr/bin/env python
"""Roofline of the streaming kernels that are NOT the bench headline (SURVEY 8d "algorithmic
bytes, others"): the single-agent env steps, tag_gridworld, the stand-alone categorical
sampler.  At the BASELINE sizes (10 000 envs) these launches move 0.4-0.6 MB and are
launch-latency-bound, so each kernel is timed twice: at its BASELINE size (time per launch)
and at a size whose working set exceeds the 126 MB L2 several times (achieved GB/s against
MEASURED_PEAKS.json `hbm_gbs`).  CUDA events on the launching stream, warm-up first, the
large inputs themselves defeat the L2.  Writes one JSON object per kernel to stdout.

    python scripts/bench_small_kernels.py > gpurun_out/small_kernels.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from warp_drive_b200 import lib as wlib  # noqa: E402

PEAK = 6576.4
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fp:
        PEAK = float(json.load(fp)["hbm_gbs"])
except (OSError, KeyError, ValueError):
    pass


def timed(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    start = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    stop = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        start[i].record()
        fn()
        stop[i].record()
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in zip(start, stop))
    return ms[len(ms) // 2], ms[0]


def report(name, unit_bytes, units_small, units_large, make):
    out = {"kernel": name, "algorithmic_bytes_per_unit": unit_bytes}
    for tag, n in (("baseline_size", units_small), ("large", units_large)):
        fn = make(n)
        med, best = timed(fn)
        out[tag] = {"units": n, "ms_median": med, "ms_min": best,
                    "GBps": unit_bytes * n / (med * 1e-3) / 1e9,
                    "frac_of_hbm_peak": unit_bytes * n / (med * 1e-3) / 1e9 / PEAK}
    print(json.dumps(out))
    sys.stdout.flush()


def single_agent(name, sdim, odim, continuous, consts, low, high):
    L, p = wlib.load(), wlib.ptr
    fn = getattr(L, f"wdb_{name}_step")

    def make(E):
        g = torch.Generator(device="cuda").manual_seed(0)
        span = torch.tensor(high, device="cuda") - torch.tensor(low, device="cuda")
        state0 = torch.rand((E, 1, sdim), device="cuda", generator=g) * span + torch.tensor(
            low, device="cuda")
        state = state0.clone()
        action = (torch.rand((E, 1, 1), device="cuda", generator=g) * 2 - 1 if continuous
                  else torch.randint(0, 2, (E, 1, 1), device="cuda", generator=g,
                                     dtype=torch.int32))
        done = torch.zeros(E, dtype=torch.int32, device="cuda")
        reward = torch.zeros((E, 1), device="cuda")
        obs = torch.zeros((E, 1, odim), device="cuda")
        ts = torch.zeros(E, dtype=torch.int32, device="cuda")

        def run():
            wlib.check(fn(wlib.stream_ptr(), E, p(state), p(action), p(done), p(reward), p(obs),
                          *consts, p(ts), 1 << 30))

        return run

    # state in + state out + obs out + action + reward
    unit = 4 * (2 * sdim + odim + 2)
    report(f"wdb_{name}_step", unit, 10000, 1 << 22, make)


def gridworld():
    L, p = wlib.load(), wlib.ptr
    N = 5

    def make(E):
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randint(0, 10, (E, N), device="cuda", generator=g, dtype=torch.int32)
        y = torch.randint(0, 10, (E, N), device="cuda", generator=g, dtype=torch.int32)
        act = torch.randint(0, 5, (E, N, 1), device="cuda", generator=g, dtype=torch.int32)
        done = torch.zeros(E, dtype=torch.int32, device="cuda")
        rew = torch.zeros((E, N), device="cuda")
        obs = torch.zeros((E, N, 4 * N + 1), device="cuda")
        ts = torch.zeros(E, dtype=torch.int32, device="cuda")
        moves = torch.tensor([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1]], dtype=torch.int32,
                             device="cuda")

        def run():
            wlib.check(L.wdb_tag_gridworld_step(
                wlib.stream_ptr(), E, N, p(x), p(y), p(act), p(done), p(rew), p(obs), 0.1, 10.0,
                5.0, 0.01, 1, 9, p(ts), 1 << 30, p(moves)))

        return run

    report("wdb_tag_gridworld_step (5 agents, full obs)", 108, 2000 * N, (1 << 20) * N,
           lambda units: make(units // N))


def sampler():
    L, p = wlib.load(), wlib.ptr
    N, A = 105, 21

    def make(units):
        E = units // N
        probs = torch.softmax(torch.randn((E, N, A), device="cuda"), -1)
        actions = torch.zeros((E, N, 1), dtype=torch.int32, device="cuda")
        rng = torch.zeros(int(L.wdb_rng_state_bytes(E * N)), dtype=torch.uint8, device="cuda")
        wlib.check(L.wdb_rng_init(wlib.stream_ptr(), p(rng), E * N, 7))

        def run():
            wlib.check(L.wdb_sample_actions(wlib.stream_ptr(), p(rng), p(probs), p(actions), None,
                                            E, N, A, 0, None, 0, 0, None))

        return run

    report("wdb_sample_actions (21 actions, no cum_distr)", 4 * A + 4, 2000 * N, 40000 * N, make)


def main():
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    single_agent("cartpole", 4, 4, False,
                 [9.8, 0.1, 1.1, 0.5, 0.05, 10.0, 0.02, 12 * 2 * np.pi / 360, 2.4],
                 [-2.0, -2.0, -0.2, -2.0], [2.0, 2.0, 0.2, 2.0])
    single_agent("mountain_car", 2, 2, False, [-1.2, 0.6, 0.07, 0.5, 0.0, 0.001, 0.0025],
                 [-1.2, -0.07], [0.6, 0.07])
    single_agent("continuous_mountain_car", 2, 2, True,
                 [-1.0, 1.0, -1.2, 0.6, 0.07, 0.45, 0.0, 0.0015], [-1.2, -0.07], [0.6, 0.07])
    single_agent("pendulum", 2, 3, True, [], [-3.0, -8.0], [3.0, 8.0])
    single_agent("acrobot", 4, 6, False, [], [-3.0, -3.0, -6.0, -12.0], [3.0, 3.0, 6.0, 12.0])
    gridworld()
    sampler()


if __name__ == "__main__":
    main()
