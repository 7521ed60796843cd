"""Custom env plug-in on the GPU: tests/custom_env/counter_env_step.cu is
compiled at run time for sm_100a, found through the env registrar, launched by name with the
reference's pycuda-style call `f(*args, block=, grid=)`, and must reproduce the NumPy twin
bit for bit (integer dynamics) through step, done-masked reset and the custom reset kernel."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
COUNTER_CU = os.path.join(HERE, "custom_env", "counter_env_step.cu")


def _wrapper(n_envs, n_agents, blocks_per_env=1):
    import sys

    sys.path.insert(0, os.path.join(HERE, "custom_env"))
    from counter_env import CUDACounterEnv, STEP_TABLE

    from warp_drive_b200.env_wrapper import EnvWrapper
    from warp_drive_b200.training.utils.data_loader import create_and_push_data_placeholders
    from warp_drive_b200.utils.env_registrar import EnvironmentRegistrar

    reg = EnvironmentRegistrar()
    reg.add_cuda_env_src_path("CounterEnv", COUNTER_CU)
    env = CUDACounterEnv(num_agents=n_agents, episode_length=20, limit=12 * n_agents,
                         env_backend="pycuda")
    wrapper = EnvWrapper(env, num_envs=n_envs, env_backend="pycuda", env_registrar=reg,
                         blocks_per_env=blocks_per_env)
    fm, dm = wrapper.cuda_function_manager, wrapper.cuda_data_manager
    assert len(fm._custom_modules) == 1 and fm._custom_modules[0].has("CudaCounterEnvStep")
    dm.add_shared_constants({"kStepTable": STEP_TABLE})
    fm.initialize_shared_constants(dm, ["kStepTable"])
    wrapper.reset_all_envs()
    create_and_push_data_placeholders(env_wrapper=wrapper, action_sampler=None,
                                      push_data_batch_placeholders=False)
    return wrapper, STEP_TABLE


@pytest.mark.parametrize("n_envs,n_agents,bpe", [(64, 5, 1), (33, 70, 1), (16, 24, 2)])
def test_custom_env_matches_its_numpy_twin(n_envs, n_agents, bpe):
    wrapper, table = _wrapper(n_envs, n_agents, bpe)
    dm = wrapper.cuda_data_manager
    actions = dm.data_on_device_via_torch("sampled_actions")
    rs = np.random.RandomState(n_agents)
    counters = np.zeros((n_envs, n_agents), np.int64)
    ts = np.zeros(n_envs, np.int64)
    limit, ep_len = 12 * n_agents, 20
    resets = 0
    for _ in range(45):
        a = rs.randint(0, 4, (n_envs, n_agents, 1)).astype(np.int32)
        actions[:] = torch.from_numpy(a).cuda()
        wrapper.step_all_envs()
        counters += table[a[..., 0]]
        ts += 1
        done = (ts == ep_len) | (counters.sum(1) >= limit)
        assert np.array_equal(dm.pull_data_from_device("counters"), counters)
        assert np.array_equal(dm.pull_data_from_device("rewards"), counters.astype(np.float32))
        obs = dm.pull_data_from_device("observations")
        assert np.array_equal(obs[..., 0], counters.astype(np.float32))
        assert np.array_equal(obs[..., 1], np.broadcast_to(counters.sum(1, keepdims=True),
                                                           counters.shape).astype(np.float32))
        assert np.array_equal(obs[..., 2], np.broadcast_to(
            (ts.astype(np.float32) / np.float32(ep_len))[:, None], counters.shape))
        assert np.array_equal(dm.pull_data_from_device("_done_").astype(bool), done)
        wrapper.reset_only_done_envs()
        counters[done] = 0
        ts[done] = 0
        resets += int(done.sum())
        assert np.array_equal(dm.pull_data_from_device("counters"), counters)
        assert not dm.pull_data_from_device("_done_").any()
    assert resets >= n_envs           # every env finished at least once on average
    assert wrapper.cuda_function_manager._custom_modules[0].launches == 45


def test_custom_reset_kernel_and_source_string():
    wrapper, _ = _wrapper(8, 5)
    dm = wrapper.cuda_data_manager
    # Cuda<Env>Reset from the same user file was registered by EnvWrapper
    wrapper.custom_reset_all_envs(args=["counters", "reset_value"])
    torch.cuda.synchronize()
    assert (dm.pull_data_from_device("counters") == 7).all()
    # a second module from a source STRING lives next to it
    fm = wrapper.cuda_function_manager
    fm.load_cuda_from_source_code(
        'extern "C" __global__ void CudaScaleCounters(int *c, int k) {\n'
        '  if (wdb_env::agent_valid()) c[wdb_env::agent_index()] *= k; }\n',
        default_functions_included=False)
    fm.initialize_functions(["CudaScaleCounters"])
    fm.get_function("CudaScaleCounters")(dm.device_data("counters"), np.int32(3),
                                         block=fm.block, grid=fm.grid)
    torch.cuda.synchronize()
    assert (dm.pull_data_from_device("counters") == 21).all()
