"""The oracle itself is pinned here, on CPU, before any GPU parity test trusts it.

  * gridworld: the reference's own literal golden vectors
    (/root/reference/tests/example_envs/pycuda_tests/test_tag_gridworld_step_cuda.py:
    132-708, extracted into tests/golden/gridworld_cuda_golden.npz) -- bit-level.
  * gridworld / tag_continuous: trajectories recorded from the reference's NumPy envs
    (tests/golden/make_golden.py) -- the reference's own CPU<->CUDA agreement rule
    (env_cpu_gpu_consistency_checker.py:543-561) with a much tighter tolerance.
  * Philox4x32-10: Random123 known-answer vectors.
"""
import numpy as np
import pytest

import oracle
from conftest import load_golden
from helpers import close, tc_cfg_from_fixture, tc_state_from_fixture

MOVES = np.array([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1]], np.int32).reshape(-1)


def _gw_oracle_step(L, x, y, actions, done, ts, rewards, obs, cfg, full, B, T):
    L.wd_oracle_tag_gridworld_step(
        x.shape[0], x.shape[1], x, y, actions, done, rewards, obs,
        float(cfg["wall_hit_penalty"]), float(cfg["tag_reward_for_tagger"]),
        float(cfg["tag_penalty_for_runner"]), float(cfg["step_cost_for_tagger"]),
        int(full), int(B), ts, int(T), MOVES)


def test_gridworld_reference_cuda_golden_vectors(oracle_lib):
    g = load_golden("gridworld_cuda_golden.npz")
    x0, y0 = g["init_loc_x"].astype(np.int32), g["init_loc_y"].astype(np.int32)
    assert (g["kIndexToActionArr"].reshape(-1) == MOVES).all()
    x, y = x0.copy(), y0.copy()
    done = np.zeros(2, np.int32)
    ts = np.zeros(2, np.int32)
    rewards = np.zeros((2, 5), np.float32)
    obs = np.zeros((2, 5, 21), np.float32)
    B = int(g["world_boundary"])
    for step in (1, 2):
        # the reference test resets done envs first (episode_length == 1), then samples
        # from one-hot distributions with the real sampler, then steps
        oracle_lib.wd_oracle_reset_when_done(x.ctypes.data, x0.ctypes.data, done, 2, 5, 0)
        oracle_lib.wd_oracle_reset_when_done(y.ctypes.data, y0.ctypes.data, done, 2, 5, 0)
        oracle_lib.wd_oracle_undo_done_and_reset_timestep(done, ts, 2, 0)
        probs = np.ascontiguousarray(g[f"agent_distribution_step{step}"], np.float32)
        actions = np.zeros((2, 5), np.int32)
        u = np.random.RandomState(step).uniform(1e-6, 1.0, 10).astype(np.float32)
        oracle_lib.wd_oracle_sample_actions(probs.reshape(10, 5), actions.reshape(-1), 1,
                                            None, u, 10, 5, 0)
        assert (actions == g[f"ref_actions_step{step}"]).all()
        _gw_oracle_step(oracle_lib, x, y, actions, done, ts, rewards, obs, g, True, B, 1)
        assert np.abs(rewards - g[f"ref_rewards_step{step}"]).max() < 1e-5
        # the reference compares obs * 4 (test_tag_gridworld_step_cuda.py:411-417)
        assert np.abs(obs * 4 - g[f"ref_observations_step{step}"]).max() < 1e-5
        assert list(done) == [1, 1]          # :419 and :708


@pytest.mark.parametrize("name", ["test1", "test2", "config1"])
def test_gridworld_vs_reference_numpy_trajectory(oracle_lib, name):
    fx = load_golden(f"gridworld_numpy_{name}.npz")
    cfg = {k[5:]: fx[k] for k in fx if k.startswith("cfg__")}
    full = bool(cfg.get("use_full_observation", True))
    B, T = int(cfg["grid_length"]), int(cfg["episode_length"])
    x0 = fx["init__loc_x"].astype(np.int32)[None].copy()
    y0 = fx["init__loc_y"].astype(np.int32)[None].copy()
    x, y = x0.copy(), y0.copy()
    done, ts = np.zeros(1, np.int32), np.zeros(1, np.int32)
    N = x.shape[1]
    rewards = np.zeros((1, N), np.float32)
    obs = np.zeros((1, N, fx["obs"].shape[2]), np.float32)
    for t in range(fx["actions"].shape[0]):
        if done[0]:
            x[:], y[:] = x0, y0
            done[0], ts[0] = 0, 0
        actions = fx["actions"][t].astype(np.int32)[None].copy()
        _gw_oracle_step(oracle_lib, x, y, actions, done, ts, rewards, obs, cfg, full, B, T)
        assert (x[0] == fx["loc_x"][t]).all() and (y[0] == fx["loc_y"][t]).all()
        assert close(rewards[0], fx["rewards"][t], 1e-6).all()
        assert close(obs[0], fx["obs"][t], 1e-6).all()
        assert bool(done[0]) == bool(fx["done"][t])


@pytest.mark.parametrize(
    "name", ["test1", "test2", "test3", "test4", "partial_mid", "config2_short"])
def test_tag_continuous_vs_reference_numpy_trajectory(oracle_lib, name):
    """Free-running oracle (CUDA semantics, float32) vs the reference NumPy env (mixed
    float64).  Discrete outcomes must agree exactly; floats to 2e-5 abs-or-rel."""
    fx = load_golden(f"tag_continuous_numpy_{name}.npz")
    cfg = tc_cfg_from_fixture(fx)
    st = tc_state_from_fixture(fx, 1)
    T = fx["actions"].shape[0]
    last = int(fx["episode_length"])
    for t in range(T):
        obs, rew = oracle.tag_continuous_step(st, cfg, fx["actions"][t][None])
        assert (st["still_in_the_game"][0] == fx["still_in_the_game"][t]).all(), t
        for key in ("loc_x", "loc_y", "speed", "direction", "acceleration"):
            assert close(st[key][0], fx[key][t], 2e-5).all(), (key, t)
        ok = close(obs[0], fx["obs"][t], 2e-5)
        assert ok.all(), (t, np.argwhere(~ok)[:5])
        if t + 1 == last:
            # known CUDA != NumPy edge: a runner tagged on the very last step also gets
            # the end-of-game bonus in the CUDA kernel
            continue
        assert close(rew[0], fx["rewards"][t], 2e-5).all(), t
        assert bool(st["_done_"][0]) == bool(fx["done"][t])


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2,
         [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert oracle.philox4x32_10(ctr, key).tolist() == want


def test_sampler_index_selection(oracle_lib):
    """search_index semantics of random.cu:33-49 on hand-checkable cases."""
    probs = np.array([[0.25, 0.25, 0.25, 0.25],
                      [0.0, 0.0, 1.0, 0.0],
                      [0.1, 0.2, 0.3, 0.4]], np.float32)
    out = np.zeros(3, np.int32)
    cum = np.zeros((3, 4), np.float32)
    for u, want in [([0.1, 0.5, 0.05], [0, 2, 0]),
                    ([0.25, 1.0, 0.1], [0, 2, 0]),       # |cdf - p| < 1e-8 shortcut
                    ([0.26, 1e-6, 0.31], [1, 2, 2]),
                    ([1.0, 0.999, 1.0], [3, 2, 3])]:
        oracle_lib.wd_oracle_sample_actions(probs, out, 1, cum, np.array(u, np.float32),
                                            3, 4, 0)
        assert out.tolist() == want, (u, out)
    assert np.allclose(cum[2], [0.1, 0.3, 0.6, 1.0], atol=1e-6)
    # CDF that never reaches p (sum < 1): clamps to the last index
    short = np.array([[0.2, 0.2, 0.2]], np.float32)
    o = np.zeros(1, np.int32)
    oracle_lib.wd_oracle_sample_actions(short, o, 1, None, np.array([0.9], np.float32),
                                        1, 3, 0)
    assert o[0] == 2
    # argmax: first maximum wins
    am = np.array([[0.3, 0.3, 0.2], [0.1, 0.5, 0.5]], np.float32)
    o = np.zeros(2, np.int32)
    oracle_lib.wd_oracle_sample_actions(am, o, 1, None, np.zeros(2, np.float32), 2, 3, 1)
    assert o.tolist() == [0, 1]


@pytest.mark.parametrize("name", ["test1", "test2", "test3", "test4", "partial_mid"])
def test_numpy_port_matches_reference(name):
    """oracle/numpy_ref.py (the reported CPU baseline) reproduces trajectories recorded
    from the real reference NumPy env, including rewards on the final step."""
    from oracle.numpy_ref import TagContinuousNumpyRef

    fx = load_golden(f"tag_continuous_numpy_{name}.npz")
    cfg = tc_cfg_from_fixture(fx)
    init = {k: fx[f"init__{k}"] for k in ("loc_x", "loc_y", "speed", "direction", "acceleration")}
    env = TagContinuousNumpyRef(cfg, init)
    N = env.N
    o0 = env.generate_observation()
    assert close(np.stack([o0[a] for a in range(N)]), fx["obs0"], 1e-6).all()
    for t in range(fx["actions"].shape[0]):
        obs, rew, done = env.step(fx["actions"][t])
        assert close(np.stack([obs[a] for a in range(N)]), fx["obs"][t], 1e-6).all(), t
        assert close(np.array([rew[a] for a in range(N)]), fx["rewards"][t], 1e-6).all(), t
        assert bool(done) == bool(fx["done"][t])
        assert (env.still_in_the_game == fx["still_in_the_game"][t]).all()


# ------------------------------------------------------- classic control
_CC_ENVS = ["cartpole", "mountain_car", "continuous_mountain_car", "pendulum", "acrobot"]


@pytest.mark.parametrize("name", _CC_ENVS)
def test_classic_control_oracle_vs_reference_kernels_in_the_cuda_simulator(name, oracle_lib):
    """tests/golden/classic_control_cudasim.npz holds inputs/outputs of the reference's OWN
    numba kernels executed by numba's CUDA simulator (tests/golden/
    make_classic_control_golden.py).  The simulator applies NumPy scalar typing instead of
    numba's compiled typing, so this pins the algorithm -- every branch, clip, wrap, reward
    and done rule, and the argument order -- to 1e-5 abs-or-rel (Acrobot 2e-4), not the last
    bit."""
    g = load_golden("classic_control_cudasim.npz")
    T = int(g["episode_length"])
    consts = [float(c) for c in g[f"{name}__consts"]]
    fn = getattr(oracle_lib, f"wd_oracle_{name}_step")
    steps, E = g[f"{name}__state_in"].shape[:2]
    flips = 0
    seen = set()
    for s in range(steps):
        state = g[f"{name}__state_in"][s].copy()
        ts = g[f"{name}__timestep_in"][s].copy()
        action = np.ascontiguousarray(g[f"{name}__action"][s])
        done = np.zeros(E, np.int32)
        reward = np.full((E, 1), 7.0, np.float32)
        obs = np.full_like(g[f"{name}__obs"][s], 7.0)
        fn(E, state, action, done, reward, obs, *consts, ts, T)
        mine = np.concatenate([state.reshape(E, -1), obs.reshape(E, -1), reward], 1)
        ref = np.concatenate([g[f"{name}__state_out"][s].reshape(E, -1),
                              g[f"{name}__obs"][s].reshape(E, -1), g[f"{name}__reward"][s]], 1)
        # Acrobot: velocities reach 28 rad/s and the RK4 stages cancel, so the simulator's
        # NumPy typing shows up at the 1e-5..1e-4 level there
        tol = 2e-4 if name == "acrobot" else 1e-5
        ok = (np.abs(mine - ref) <= tol + tol * np.abs(ref)).all(1)
        flips += int((~ok).sum()) + int((done[ok] != g[f"{name}__done"][s][ok]).sum())
        assert (ts == g[f"{name}__timestep_out"][s]).all()
        seen |= set(np.unique(done).tolist())
    assert flips == 0, (name, flips)
    assert 1 in seen and (name != "mountain_car" or 2 in seen)


@pytest.mark.parametrize("name,cls", [
    ("mountain_car", "MountainCarPhysics"),
    ("continuous_mountain_car", "ContinuousMountainCarPhysics"),
    ("pendulum", "PendulumPhysics"), ("acrobot", "AcrobotPhysics")])
def test_classic_control_oracle_vs_float64_physics(name, cls, oracle_lib):
    """The C oracle (the reference's numba kernels restated) against the float64 restatement
    of the gym integrators that the CPU envs use: same physics, different precision."""
    from warp_drive_b200.envs.single_agent import classic_control as cc

    phys = getattr(cc, cls)()
    g = load_golden("classic_control_cudasim.npz")
    consts = [float(c) for c in g[f"{name}__consts"]]
    fn = getattr(oracle_lib, f"wd_oracle_{name}_step")
    state_in = g[f"{name}__state_in"][0]
    action = np.ascontiguousarray(g[f"{name}__action"][0])
    E = state_in.shape[0]
    state = state_in.copy()
    done = np.zeros(E, np.int32)
    reward = np.zeros((E, 1), np.float32)
    obs = np.zeros_like(g[f"{name}__obs"][0])
    ts = np.zeros(E, np.int32)
    fn(E, state, action, done, reward, obs, *consts, ts, 1000)
    bad = 0
    for e in range(E):
        phys.state = np.array(state_in[e, 0], dtype=np.float64)
        o, r, terminated, _, _ = phys.step(action[e, 0, 0] if action.dtype.kind == "i"
                                           else action[e, 0])
        ok = np.allclose(obs[e, 0], o, rtol=1e-4, atol=1e-5) and abs(reward[e, 0] - r) <= 1e-4 * max(1, abs(r))
        bad += (not ok) or (bool(done[e]) != bool(terminated))
    assert bad <= 1, (name, bad)     # one row may sit on a wrap / clip / goal threshold
