#!/usr/bin/env python
"""Interleaved A/B of the rollout-step launch modes at BASELINE config 2, in ONE process:
  fork       one forward launch per policy on two streams (fork / join by events): the default
  fork+tail  the same, env kernel with the last partial wave as one-env CTAs (tc_tail_split)
  fork+pdlenv  the same, env step launched programmatically dependent on the runner forward
  pair       both policies' forwards in one launch, plain stream order
  pair+pdl   ... plus programmatic dependent launches
(AB_MODES=a,b restricts the set.)
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
    # name -> (both policies in one launch?, library options while this engine's graph is captured)
    modes = {"fork+pdlenv": (False, {"pdl": 1, "tc_tail_split": 0}),
             "fork+tail": (False, {"pdl": 0, "tc_tail_split": 1}),
             "fork": (False, {"pdl": 0, "tc_tail_split": 0}),
             "pair": (True, {"pdl": 0, "tc_tail_split": 0}),
             "pair+pdl": (True, {"pdl": 1, "tc_tail_split": 0})}
    only = os.environ.get("AB_MODES")
    if only:
        modes = {k: v for k, v in modes.items() if k in only.split(",")}
    defaults = {"pdl": 0, "tc_tail_split": 0}
    engines = {}
    for name, (pair, opts) in modes.items():
        for k, v in opts.items():
            assert L.wdb_set_option(k.encode(), v) == 0
        _w, eng, _s, _pm = bench.build_engine(E, seed=1234, graph_steps=T, pair_forward=pair)
        eng.pdl_after_fork = name == "fork+pdlenv"
        for _ in range(3):          # capture (with this mode's options) + warm replays
            eng.rollout()
        torch.cuda.synchronize()
        engines[name] = (eng, _w)
    for k, v in defaults.items():
        L.wdb_set_option(k.encode(), v)
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
