#!/usr/bin/env python
"""Interleaved A/B of the rollout-step launch modes at BASELINE config 2, in ONE process:
  pair+pdl   both policies' forwards in one launch, programmatic dependent launches
  pair       one launch, plain stream order
  fork       one launch per policy on two streams (fork / join by events)
Each mode gets its own engine (same seed) and CUDA graph of T steps; the modes are timed
round-robin (CUDA events around `reps` replays) so clock / thermal drift hits all alike.
Prints one JSON object: per mode the per-round ms per step and the median."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from warp_drive_b200 import lib as wlib  # noqa: E402


def main():
    E = int(os.environ.get("AB_ENVS", 2000))
    T = int(os.environ.get("AB_T", 20))
    rounds = int(os.environ.get("AB_ROUNDS", 7))
    reps = int(os.environ.get("AB_REPS", 10))
    L = wlib.load()
    modes = {"pair+pdl": (True, 1), "pair": (True, 0), "fork": (False, 0)}
    engines = {}
    for name, (pair, pdl) in modes.items():
        assert L.wdb_set_option(b"pdl", pdl) == 0
        _w, eng, _s, _pm = bench.build_engine(E, seed=1234, graph_steps=T, pair_forward=pair)
        for _ in range(3):          # capture (with this mode's pdl setting) + warm replays
            eng.rollout()
        torch.cuda.synchronize()
        engines[name] = (eng, _w)
    L.wdb_set_option(b"pdl", 0)
    out = {k: [] for k in modes}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(rounds):
        for name, (eng, _w) in engines.items():
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                eng.rollout()
            b.record()
            torch.cuda.synchronize()
            out[name].append(a.elapsed_time(b) / (reps * T))
    res = {k: {"ms_per_step_rounds": [round(x, 5) for x in v],
               "median": sorted(v)[len(v) // 2], "min": min(v)} for k, v in out.items()}
    print(json.dumps({"envs": E, "T": T, "reps": reps, "modes": res}))


if __name__ == "__main__":
    main()
