"""Same-box GPU anchor: the REFERENCE's own CUDA kernels (oracle/_ref/*.fatbin, compiled
from /root/reference in place for sm_100a) timed on this GPU.

TEST / MEASUREMENT INFRASTRUCTURE -- never on the product path.  bench.py calls this
outside its timed region and prints the result as `ref_gpu` next to the product numbers
(BASELINE.md section 2 "B-REF-GPU" / "B-REF-E2E").

What the reference launches for the env side of ONE rollout timestep of tag_continuous
(warp_drive/training/trainers/trainer_base.py:383-428):
  * sample_actions x n_heads         (cuda_includes/core/random.cu:51-85)
  * CudaTagContinuousStep            (tag_continuous_step_pycuda.cu:351-520)
  * host sync for `done_flags.any()` (trainer_base.py:421-422)
  * if any env is done: one reset_in_*_when_done_* launch per registered array
    (reset.cu:9-75, pycuda_function_manager.py:668-753) + undo_done_flag_and_reset_timestep
It also synchronises after the sampling phase and after the step phase (trainer_base.py:
396-427).  `sequence_*` below replays exactly that pattern; `step_kernel_us` is the step
kernel alone with the L2 flushed before every launch (the product's roofline leg uses the
same protocol).
"""
import numpy as np
import torch

from . import ref_cuda


def _median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def time_reference_tag_continuous(env, n_envs, iters=20, flush=None, seed=7, bpe=1):
    """env: a reset() TagContinuous host object (product env class: only used to read the
    initial state and constants).  Returns a dict of microsecond timings or
    {"unavailable": why}."""
    N = env.num_agents
    if not ref_cuda.available(n_envs, N, bpe):
        return {"unavailable": f"oracle/_ref/ref_E{n_envs}_N{N}_B{bpe}.fatbin not built"}
    dev = "cuda"
    dd = env.get_data_dictionary()
    K = int(env.num_other_agents_observed)
    F = 7 * K + 1
    E = n_envs
    ref = ref_cuda.RefModule(E, N, bpe)

    def rep(a, dt):
        a = np.asarray(a, dt)
        return torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, (E,) + a.shape))).to(dev)

    st = {k: rep(dd[k]["data"], np.float32) for k in
          ("loc_x", "loc_y", "speed", "direction", "acceleration", "edge_hit_reward_penalty")}
    st["still_in_the_game"] = rep(dd["still_in_the_game"]["data"], np.int32)
    st["num_runners"] = torch.full((E,), env.num_runners, dtype=torch.int32, device=dev)
    st["nearest_neighbor_ids"] = torch.zeros((E, N, K), dtype=torch.int32, device=dev)
    st["_done_"] = torch.zeros(E, dtype=torch.int32, device=dev)
    st["_timestep_"] = torch.zeros(E, dtype=torch.int32, device=dev)
    cfg = {"agent_types": torch.as_tensor(np.asarray(dd["agent_types"]["data"], np.int32)).to(dev)}
    for k in ("acceleration_actions", "turn_actions", "skill_levels", "step_rewards"):
        cfg[k] = torch.as_tensor(np.asarray(dd[k]["data"], np.float32)).to(dev)
    for k in ("grid_length", "edge_hit_penalty", "max_speed", "distance_margin_for_reward",
              "tag_reward_for_tagger", "tag_penalty_for_runner",
              "end_of_game_reward_for_runner", "num_other_agents_observed",
              "use_full_observation", "runner_exits_game_after_tagged"):
        cfg[k] = dd[k]["data"]
    cfg["episode_length"] = env.episode_length
    nd = torch.zeros((E, N, N - 1), device=dev)
    nid = torch.zeros((E, N, N - 1), dtype=torch.int32, device=dev)
    obs = torch.zeros((E, N, F), device=dev)
    rew = torch.zeros((E, N), device=dev)
    actions = torch.zeros((E, N, 2), dtype=torch.int32, device=dev)
    head = [torch.zeros((E, N, 1), dtype=torch.int32, device=dev) for _ in range(2)]
    A = len(np.asarray(dd["acceleration_actions"]["data"]))
    probs = [torch.softmax(torch.randn((E, N, A), device=dev), -1) for _ in range(2)]
    cum = [torch.zeros((E, N, A), device=dev) for _ in range(2)]
    # the arrays the reference resets (save_copy_and_apply_at_reset=True in
    # tag_continuous.py:680-756 + `observations` from the trainer's data loader)
    reset_list = [(st[k], st[k].clone(), tuple(st[k].shape)) for k in
                  ("loc_x", "loc_y", "speed", "direction", "acceleration", "num_runners",
                   "edge_hit_reward_penalty", "nearest_neighbor_ids", "still_in_the_game")]
    reset_list += [(nd, nd.clone(), tuple(nd.shape)), (nid, nid.clone(), tuple(nid.shape)),
                   (obs, obs.clone(), tuple(obs.shape))]
    ref.init_random(seed)
    torch.cuda.synchronize()

    def sample():
        for k in range(2):
            ref.sample_actions(probs[k], head[k], cum[k], N, A)
        # the reference assembles [E, N, 2] from the per-head arrays with a torch copy
        actions[..., 0:1].copy_(head[0])
        actions[..., 1:2].copy_(head[1])

    def step():
        ref.tag_continuous_step(st, cfg, actions, obs, rew, nd, nid)

    def reset(force=0):
        for data, at_reset, shape in reset_list:
            ref.reset_when_done(data, at_reset, st["_done_"], shape, force)
        ref.undo_done(st["_done_"], st["_timestep_"], force)

    if flush is None:
        flush = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    # ---- kernels alone, L2 flushed before every launch
    t_step, t_sample, t_reset = [], [], []
    for i in range(iters + 2):
        sample()
        flush.add_(1)
        a, b = ev(), ev()
        a.record(); step(); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            t_step.append(a.elapsed_time(b) * 1e3)
        reset(0)
        flush.add_(1)
        a, b = ev(), ev()
        a.record(); ref.sample_actions(probs[0], head[0], cum[0], N, A); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            t_sample.append(a.elapsed_time(b) * 1e3)
    # all 12 + 1 reset launches, no env done (the masked no-op cost) and all envs forced
    for force in (0, 1):
        ts = []
        for i in range(5):
            flush.add_(1)
            a, b = ev(), ev()
            a.record(); reset(force); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        t_reset.append(_median(ts))
    reset(1)

    # ---- the reference's per-timestep launch / sync pattern (env side, no policy forward)
    def sequence(with_reset):
        sample()
        torch.cuda.synchronize()                # trainer_base.py:404-406
        step()
        any_done = bool(st["_done_"].any())     # trainer_base.py:421 (host sync)
        if any_done or with_reset:
            reset(0)
        torch.cuda.synchronize()                # trainer_base.py:424-426

    out = {}
    for key, with_reset in (("sequence_us_reset_only_when_done", False),
                            ("sequence_us_reset_every_step", True)):
        reset(1)
        torch.cuda.synchronize()
        ts = []
        for i in range(iters + 2):
            a, b = ev(), ev()
            a.record(); sequence(with_reset); b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b) * 1e3)
        out[key] = _median(ts)
    out.update({
        "kernels": f"oracle/_ref/ref_E{E}_N{N}_B{bpe}.fatbin = the reference's CUDA-C sources "
                   "compiled in place with nvcc --fatbin -arch=sm_100a, launched with the "
                   f"reference geometry (grid {E * bpe}, block {(N - 1) // bpe + 1})",
        "step_kernel_us": _median(t_step), "step_kernel_us_min": min(t_step),
        "sample_actions_kernel_us": _median(t_sample),
        "reset_13_launches_us_none_done": t_reset[0],
        "reset_13_launches_us_all_done": t_reset[1],
        "protocol": "CUDA events, L2 flushed (512 MiB read+write) before every launch for "
                    "the *_kernel_us numbers; sequence_* = 2 x sample_actions + 2 torch "
                    "copies + CudaTagContinuousStep + done.any() host sync (+ 13 reset "
                    "launches) with the reference's synchronize() calls, no flush",
        "envs": E, "agents": N,
    })
    return out
