"""GPU tests of the core services through the public managers / C ABI:
RNG, categorical + OU samplers, fused reset (+ pools), episode log, testkernel.

Expected values marked (ref ...) are the known answers asserted by the reference's own
manager tests (/root/reference/tests/warp_drive/pycuda_tests/*.py)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _managers(num_agents=5, num_envs=2, episode_length=2):
    from warp_drive_b200.managers.data_manager import CUDADataManager
    from warp_drive_b200.managers.function_manager import CUDAFunctionManager

    dm = CUDADataManager(num_agents=num_agents, num_envs=num_envs,
                         episode_length=episode_length)
    fm = CUDAFunctionManager(num_agents=num_agents, num_envs=num_envs)
    fm.load_cuda_from_binary_file("warp_drive/cuda_bin/test_build.fatbin")  # ignored
    return dm, fm


def test_device_philox_matches_scalar_oracle(wdb_lib):
    from warp_drive_b200 import lib as wlib

    n, seed = 1000, 0x1234567890ABCDEF
    state = torch.zeros(int(wdb_lib.wdb_rng_state_bytes(n)), dtype=torch.uint8, device="cuda")
    wlib.check(wdb_lib.wdb_rng_init(wlib.stream_ptr(), state.data_ptr(), n, seed))
    out = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    for draw in range(3):
        wlib.check(wdb_lib.wdb_rng_draw_u32x4(wlib.stream_ptr(), state.data_ptr(),
                                              out.data_ptr(), n))
        got = out.cpu().numpy().view(np.uint32)
        for i in (0, 1, 17, 999):
            want = oracle.philox4x32_10([draw, 0, i, 0],
                                        [seed & 0xFFFFFFFF, seed >> 32])
            assert got[i].tolist() == want.tolist()


@pytest.mark.parametrize("n_envs,n_agents,n_actions", [(2, 5, 3), (64, 105, 21), (1000, 1, 2)])
def test_categorical_sampler_bit_exact_vs_oracle(oracle_lib, n_envs, n_agents, n_actions):
    """Same float32 CDF + search_index as random.cu:33-85, given the same uniforms."""
    from warp_drive_b200.managers.function_manager import CUDASampler
    from warp_drive_b200.utils.data_feed import DataFeed

    dm, fm = _managers(n_agents, n_envs, 5)
    sampler = CUDASampler(fm)
    sampler.init_random(seed=1)
    feed = DataFeed()
    feed.add_data(name="sampled_actions", data=np.zeros((n_envs, n_agents, 1), np.int32))
    dm.push_data_to_device(feed, torch_accessible=True)
    sampler.register_actions(dm, "sampled_actions", n_actions)
    g = torch.Generator(device="cpu").manual_seed(0)
    probs = torch.softmax(3 * torch.randn(n_envs, n_agents, n_actions, generator=g), -1)
    u = torch.rand(n_envs * n_agents, generator=g).clamp_min(1e-7)
    sampler.sample(dm, probs.cuda(), "sampled_actions", uniforms=u.cuda())
    got = dm.pull_data_from_device("sampled_actions").reshape(-1)
    got_cum = dm.pull_data_from_device("sampled_actions_cum_distr")
    want = np.zeros(n_envs * n_agents, np.int32)
    want_cum = np.zeros((n_envs * n_agents, n_actions), np.float32)
    oracle_lib.wd_oracle_sample_actions(
        probs.numpy().reshape(-1, n_actions).copy(), want, 1, want_cum, u.numpy().copy(),
        n_envs * n_agents, n_actions, 0)
    assert (got == want).all()
    assert (got_cum.reshape(-1, n_actions) == want_cum).all()   # bit-exact float32 CDF
    # argmax branch
    sampler.sample(dm, probs.cuda(), "sampled_actions", use_argmax=True)
    got = dm.pull_data_from_device("sampled_actions").reshape(-1)
    oracle_lib.wd_oracle_sample_actions(
        probs.numpy().reshape(-1, n_actions).copy(), want, 1, None, u.numpy().copy(),
        n_envs * n_agents, n_actions, 1)
    assert (got == want).all()


def test_sampler_statistics_like_reference_test():
    """ref tests/warp_drive/pycuda_tests/test_action_sampler.py:90-156: 10 000 draws,
    frequencies within 10 % of p; one-hot rows exact."""
    from warp_drive_b200.managers.function_manager import CUDASampler
    from warp_drive_b200.utils.data_feed import DataFeed

    dm, fm = _managers(5, 2, 5)
    sampler = CUDASampler(fm)
    sampler.init_random(seed=None)
    feed = DataFeed()
    feed.add_data(name="a", data=np.zeros((2, 5, 1), np.int32))
    dm.push_data_to_device(feed, torch_accessible=True)
    sampler.register_actions(dm, "a", 3)
    probs = torch.tensor([[[0.333, 0.333, 0.334]] * 5, [[0.2, 0.5, 0.3]] * 4 + [[0.0, 1.0, 0.0]]],
                         dtype=torch.float32).cuda()
    draws = []
    for _ in range(10000):
        sampler.sample(dm, probs, "a")
        draws.append(dm.data_on_device_via_torch("a").clone())
    draws = torch.stack(draws).cpu().numpy()[..., 0]          # [10000, 2, 5]
    for env in range(2):
        for agent in range(5):
            freq = np.bincount(draws[:, env, agent], minlength=3) / 10000.0
            p = probs[env, agent].cpu().numpy()
            assert np.abs(freq - p).max() < 0.1 * max(p.max(), 0.1) + 0.01
    assert (draws[:, 1, 4] == 1).all()
    # different (env, agent) streams are not identical
    assert draws[:, 0, 0].std() > 0.5 and (draws[:, 0, 0] != draws[:, 0, 1]).mean() > 0.3


def test_ou_sampler_exact_and_statistics(oracle_lib):
    """ref numba test_ou_sampler.py:64-82: stationary std = stddev / sqrt(1-(1-d)^2)."""
    from warp_drive_b200.managers.function_manager import CUDASampler
    from warp_drive_b200.utils.data_feed import DataFeed

    E, N = 2000, 1
    dm, fm = _managers(N, E, 5)
    sampler = CUDASampler(fm)
    sampler.init_random(seed=7)
    feed = DataFeed()
    feed.add_data(name="act", data=np.zeros((E, N, 1), np.float32))
    dm.push_data_to_device(feed, torch_accessible=True)
    sampler.register_actions(dm, "act", 1, is_deterministic=True)
    mean = torch.full((E, N, 1), 0.25, device="cuda")
    # exactness with injected normals
    normals = torch.randn(E * N, generator=torch.Generator().manual_seed(1))
    sampler.sample(dm, mean, "act", damping=0.15, stddev=0.2, scale=1.5, normals=normals.cuda())
    want_a = np.zeros(E * N, np.float32)
    want_s = np.zeros(E * N, np.float32)
    oracle_lib.wd_oracle_ou_process(mean.cpu().numpy().reshape(-1).copy(), want_a, want_s,
                                    normals.numpy().copy(), E * N, 0.15, 0.2, 1.5)
    assert np.allclose(dm.pull_data_from_device("act").reshape(-1), want_a, atol=1e-6)
    assert np.allclose(dm.pull_data_from_device("act_ou_state").reshape(-1), want_s, atol=1e-6)
    # scale < 1e-8 bypass
    sampler.sample(dm, mean, "act", scale=0.0)
    assert (dm.pull_data_from_device("act") == 0.25).all()
    # statistics with the device RNG
    dm.data_on_device_via_torch("act_ou_state").zero_()
    damping, stddev = 0.15, 0.2
    for _ in range(300):
        sampler.sample(dm, mean, "act", damping=damping, stddev=stddev, scale=1.0)
    ou = dm.pull_data_from_device("act_ou_state").reshape(-1)
    theory = stddev / np.sqrt(1 - (1 - damping) ** 2)
    assert abs(ou.std() - theory) < 0.05 * theory
    assert abs(ou.mean()) < 0.05


def test_function_manager_testkernel_log_and_reset():
    """Re-runs the scenario of ref tests/warp_drive/pycuda_tests/test_function_manager.py:
    69-245 and asserts the same known answers."""
    from warp_drive_b200.managers.function_manager import (
        CUDAEnvironmentReset, CUDALogController)
    from warp_drive_b200.utils.data_feed import DataFeed

    dm, fm = _managers(5, 2, 2)
    dc = CUDALogController(function_manager=fm)
    resetter = CUDAEnvironmentReset(function_manager=fm)
    fm.initialize_functions(["testkernel"])
    data = DataFeed()
    data.add_data(name="X", data=[[0.1, 0.2, 0.3, 0.4, 0.5], [0.6, 0.7, 0.8, 0.9, 1.0]],
                  save_copy_and_apply_at_reset=True, log_data_across_episode=True)
    data.add_data(name="Y", data=np.array([[6, 7, 8, 9, 10], [1, 2, 3, 4, 5]]),
                  save_copy_and_apply_at_reset=True, log_data_across_episode=True)
    data.add_data(name="multiplier", data=2.0)
    tensor = DataFeed()
    tensor.add_data(name="sampled_actions", data=np.zeros((2, 5, 3), np.int64))
    dm.push_data_to_device(data)
    dm.push_data_to_device(tensor, torch_accessible=True)
    assert dm.is_data_on_device("X") and dm.is_data_on_device_via_torch("sampled_actions")
    assert list(dm.pull_data_from_device("_log_mask_")) == [0, 0, 0]

    def dummy_step(target, step):
        resetter.reset_when_done(dm)
        fm.get_function("testkernel")(
            dm.device_data("X"), dm.device_data("Y"), dm.device_data("_done_"),
            dm.device_data("sampled_actions"), dm.device_data("multiplier"),
            np.int32(target), np.int32(step), dm.meta_info("episode_length"),
            block=fm.block, grid=fm.grid)

    dc.reset_log(data_manager=dm, env_id=0)
    for t in range(1, 3):
        dummy_step(100, t)
        dc.update_log(data_manager=dm, step=t)
    log = dc.fetch_log(data_manager=dm, names=["X", "Y"])
    assert abs(log["X_for_log"][1].mean() - 0.15) < 1e-6      # ref :128
    assert abs(log["X_for_log"][2].mean() - 0.075) < 1e-6     # ref :129
    assert log["Y_for_log"][1].mean() == 16 and log["Y_for_log"][2].mean() == 32
    actions = dm.pull_data_from_device("sampled_actions")
    for env in range(2):
        for k in range(3):
            assert actions[env, :, k].mean() == k              # ref :136-141
    assert list(dm.pull_data_from_device("_done_")) == [1, 1]  # ref :146-147
    resetter.reset_when_done(data_manager=dm)
    assert list(dm.pull_data_from_device("_done_")) == [0, 0]
    X, Y = dm.pull_data_from_device("X"), dm.pull_data_from_device("Y")
    assert abs(X[0].mean() - 0.3) < 1e-6 and abs(X[1].mean() - 0.8) < 1e-6   # ref :159-160
    assert Y[0].mean() == 8 and Y[1].mean() == 3
    dc.reset_log(data_manager=dm, env_id=1)
    assert list(dm.pull_data_from_device("_log_mask_")) == [1, 0, 0]
    log = dc.fetch_log(data_manager=dm, names=["X", "Y"])
    assert len(log["X_for_log"]) == 1 and len(log["Y_for_log"]) == 1
    for t in range(1, 3):
        dummy_step(100, t)
        dc.update_log(data_manager=dm, step=t)
    log = dc.fetch_log(data_manager=dm, names=["X", "Y"])
    assert abs(log["X_for_log"][1].mean() - 0.40) < 1e-6      # ref :189-192
    assert abs(log["X_for_log"][2].mean() - 0.20) < 1e-6
    assert log["Y_for_log"][1].mean() == 6 and log["Y_for_log"][2].mean() == 12
    dummy_step(15, 1)
    assert list(dm.pull_data_from_device("_done_")) == [1, 0]  # ref :205-206
    dummy_step(15, 2)
    assert list(dm.pull_data_from_device("_done_")) == [1, 1]  # ref :217-218
    X, Y = dm.pull_data_from_device("X"), dm.pull_data_from_device("Y")
    assert abs(X[0].mean() - 0.15) < 1e-6 and abs(X[1].mean() - 0.20) < 1e-6  # ref :227-230
    assert Y[0].mean() == 16 and Y[1].mean() == 12


@pytest.mark.parametrize("force", [False, True])
def test_fused_reset_matches_oracle(oracle_lib, force):
    """Only done envs are restored, arrays of every rank / dtype, done+timestep undone
    (ref test_env_reset.py:143-165, 224-245 + reset.cu semantics via the oracle)."""
    from warp_drive_b200.managers.function_manager import CUDAEnvironmentReset
    from warp_drive_b200.utils.data_feed import DataFeed

    E, N = 37, 7
    dm, fm = _managers(N, E, 10)
    rs = np.random.RandomState(0)
    shapes = {"a1": (E,), "a2": (E, N), "a3": (E, N, 3), "a4": (E, N, 2, 5), "i2": (E, N)}
    init = {k: (rs.randint(0, 100, s).astype(np.int32) if k[0] == "i"
                else rs.randn(*s).astype(np.float32)) for k, s in shapes.items()}
    feed = DataFeed()
    for k, v in init.items():
        feed.add_data(name=k, data=v, save_copy_and_apply_at_reset=True)
    dm.push_data_to_device(feed)
    resetter = CUDAEnvironmentReset(function_manager=fm)
    cur = {}
    for k, v in init.items():
        cur[k] = (v + 1).copy()
        dm.data_on_device_via_torch(k).copy_(torch.from_numpy(cur[k]))
    done = (rs.rand(E) < 0.3).astype(np.int32)
    ts = rs.randint(1, 10, E).astype(np.int32)
    dm.data_on_device_via_torch("_done_").copy_(torch.from_numpy(done))
    dm.data_on_device_via_torch("_timestep_").copy_(torch.from_numpy(ts))
    resetter.reset_when_done(dm, mode="force_reset" if force else "if_done")
    for k in init:
        want = cur[k].copy()
        per_env = int(np.prod(want.shape[1:])) if want.ndim > 1 else 1
        oracle_lib.wd_oracle_reset_when_done(want.ctypes.data, init[k].ctypes.data, done, E,
                                             per_env, int(force))
        assert (dm.pull_data_from_device(k) == want).all(), k
    wd, wt = done.copy(), ts.copy()
    oracle_lib.wd_oracle_undo_done_and_reset_timestep(wd, wt, E, int(force))
    assert (dm.pull_data_from_device("_done_") == wd).all()
    assert (dm.pull_data_from_device("_timestep_") == wt).all()


def test_pool_reset_statistics():
    """ref numba test_pool_reset.py:103-145: the mean over many resets approaches the
    pool mean; every restored row is a pool row."""
    from warp_drive_b200.managers.function_manager import CUDAEnvironmentReset
    from warp_drive_b200.utils.data_feed import DataFeed

    E, N, rows = 2000, 3, 10
    dm, fm = _managers(N, E, 10)
    pool = np.arange(rows * N * 2, dtype=np.float32).reshape(rows, N, 2)
    feed = DataFeed()
    feed.add_data(name="state", data=np.zeros((E, N, 2), np.float32))
    dm.push_data_to_device(feed)
    pfeed = DataFeed()
    pfeed.add_pool_for_reset(name="state_reset_pool", data=pool, reset_target="state")
    dm.push_data_to_device(pfeed)
    resetter = CUDAEnvironmentReset(function_manager=fm)
    resetter.init_reset_pool(dm, seed=3)
    dm.data_on_device_via_torch("_done_").fill_(1)
    resetter.reset_when_done(dm)
    first = dm.pull_data_from_device("state")
    ids = (first[:, 0, 0] / (N * 2)).astype(int)
    assert (first == pool[ids]).all()
    assert len(np.unique(ids)) == rows
    assert abs(first.mean() - pool.mean()) < 0.05 * pool.mean()
    dm.data_on_device_via_torch("_done_").fill_(1)
    resetter.reset_when_done(dm)
    second = dm.pull_data_from_device("state")
    assert (first != second).any()          # the per-env stream advanced
    assert list(np.unique(dm.pull_data_from_device("_done_"))) == [0]


def test_gridworld_golden_through_managers():
    """The reference's literal known-answer test, end to end through our managers:
    one-hot distributions -> real sampler -> reset_when_done -> CudaTagGridWorldStep
    (ref tests/example_envs/pycuda_tests/test_tag_gridworld_step_cuda.py:60-708)."""
    from warp_drive_b200.managers.function_manager import (
        CUDAEnvironmentReset, CUDASampler)
    from warp_drive_b200.utils.data_feed import DataFeed

    g = load_golden("gridworld_cuda_golden.npz")
    dm, fm = _managers(5, 2, 1)
    resetter = CUDAEnvironmentReset(function_manager=fm)
    sampler = CUDASampler(function_manager=fm)
    sampler.init_random(seed=None)
    fm.initialize_functions(["CudaTagGridWorldStep"])
    dm.add_shared_constants({"kIndexToActionArr": g["kIndexToActionArr"].tolist()})
    fm.initialize_shared_constants(dm, constant_names=["kIndexToActionArr"])
    data = DataFeed()
    for k in ("wall_hit_penalty", "tag_reward_for_tagger", "tag_penalty_for_runner",
              "step_cost_for_tagger"):
        data.add_data(name=k, data=float(g[k]))
    data.add_data(name="world_boundary", data=int(g["world_boundary"]))
    data.add_data(name="use_full_observation", data=True)
    data.add_data(name="loc_x", data=g["init_loc_x"], save_copy_and_apply_at_reset=True)
    data.add_data(name="loc_y", data=g["init_loc_y"], save_copy_and_apply_at_reset=True)
    dm.push_data_to_device(data)
    tensor = DataFeed()
    tensor.add_data(name="rewards", data=np.zeros((2, 5)))
    tensor.add_data(name="observations", data=np.zeros((2, 5, 21), dtype=np.float32))
    tensor.add_data(name="sampled_actions", data=np.zeros((2, 5), dtype=np.int64))
    dm.push_data_to_device(tensor, torch_accessible=True)
    sampler.register_actions(dm, action_name="sampled_actions", num_actions=5)
    for step in (1, 2):
        probs = torch.from_numpy(g[f"agent_distribution_step{step}"]).float().cuda()
        sampler.sample(dm, probs, action_name="sampled_actions")
        resetter.reset_when_done(dm)
        fm.get_function("CudaTagGridWorldStep")(
            dm.device_data("loc_x"), dm.device_data("loc_y"),
            dm.device_data("sampled_actions"), dm.device_data("_done_"),
            dm.device_data("rewards"), dm.device_data("observations"),
            dm.device_data("wall_hit_penalty"), dm.device_data("tag_reward_for_tagger"),
            dm.device_data("tag_penalty_for_runner"), dm.device_data("step_cost_for_tagger"),
            dm.device_data("use_full_observation"), dm.device_data("world_boundary"),
            dm.device_data("_timestep_"), dm.meta_info("episode_length"),
            block=fm.block, grid=fm.grid)
        assert np.abs(dm.pull_data_from_device("rewards") - g[f"ref_rewards_step{step}"]).max() < 1e-5
        obs = dm.pull_data_from_device("observations")
        assert np.abs(obs.reshape(2, 5, -1) * 4 - g[f"ref_observations_step{step}"]).max() < 1e-5
        assert (dm.pull_data_from_device("sampled_actions") == g[f"ref_actions_step{step}"]).all()
        assert list(dm.pull_data_from_device("_done_")) == [1, 1]
