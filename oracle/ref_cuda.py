"""GPU-side bit-level oracle: the REFERENCE's own CUDA kernels, run on the same device.

TEST INFRASTRUCTURE.  Loads oracle/_ref/ref_E{E}_N{N}_B{B}.fatbin (compiled from the
reference sources in place by oracle/build_ref.py) with the CUDA driver API from
cuda-python (the reference uses pycuda, which is not installed) and launches the
reference kernels on torch tensors with the reference's launch geometry
(warp_drive/managers/function_manager.py:65-67: block = ceil(N / bpe), grid = E * bpe).

Used by tests/test_parity_vs_reference_cuda.py (our kernels vs the reference kernels on
identical inputs) and by bench.py's `reference_cuda_kernel` field.
"""
import ctypes
import os

import numpy as np
import torch

try:
    from cuda.bindings import driver as cu
except ImportError:  # older cuda-python
    from cuda import cuda as cu

_HERE = os.path.dirname(os.path.abspath(__file__))

_vp, _ci, _cf, _cb = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_bool


def _ok(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


def fatbin_path(n_envs, n_agents, bpe=1):
    return os.path.join(_HERE, "_ref", f"ref_E{n_envs}_N{n_agents}_B{bpe}.fatbin")


def available(n_envs, n_agents, bpe=1):
    return os.path.exists(fatbin_path(n_envs, n_agents, bpe))


class RefModule:
    def __init__(self, n_envs, n_agents, bpe=1):
        path = fatbin_path(n_envs, n_agents, bpe)
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `python oracle/build_ref.py {n_envs} {n_agents} {bpe}` "
                "in the build container (needs /root/reference)")
        torch.zeros(1, device="cuda")  # make torch's primary context current
        self.E, self.N, self.bpe = n_envs, n_agents, bpe
        with open(path, "rb") as fp:
            self._image = fp.read()
        self.module = _ok(cu.cuModuleLoadData(self._image))
        self._fn = {}
        self.block = ((n_agents - 1) // bpe + 1, 1, 1)
        self.grid = (n_envs * bpe, 1, 1)
        self._random_ready = False

    def fn(self, name):
        if name not in self._fn:
            self._fn[name] = _ok(cu.cuModuleGetFunction(self.module, name.encode()))
        return self._fn[name]

    def launch(self, name, args, types, block=None, grid=None):
        block = block or self.block
        grid = grid or self.grid
        stream = torch.cuda.current_stream().cuda_stream
        vals = tuple(a.data_ptr() if torch.is_tensor(a) else a for a in args)
        (err,) = cu.cuLaunchKernel(self.fn(name), grid[0], grid[1], 1, block[0], 1, 1, 0,
                                   stream, (vals, tuple(types)), 0)
        if int(err) != 0:
            raise RuntimeError(f"cuLaunchKernel({name}) failed: {err}")

    def set_constant(self, name, host_array):
        dptr, size = _ok(cu.cuModuleGetGlobal(self.module, name.encode()))
        host = np.ascontiguousarray(host_array)
        assert host.nbytes <= size
        (err,) = cu.cuMemcpyHtoD(dptr, host.ctypes.data, host.nbytes)
        assert int(err) == 0

    # ------------------------------------------------------------------ kernels
    def tag_continuous_step(self, st, cfg, actions, obs, rewards, nd, nid):
        """CudaTagContinuousStep with the reference argument order
        (tag_continuous_step_pycuda.cu:351-385).  st/cfg hold torch CUDA tensors."""
        args = [
            st["loc_x"], st["loc_y"], st["speed"], st["direction"], st["acceleration"],
            cfg["agent_types"], st["edge_hit_reward_penalty"],
            float(cfg["edge_hit_penalty"]), float(cfg["grid_length"]),
            cfg["acceleration_actions"], cfg["turn_actions"], float(cfg["max_speed"]),
            int(cfg["num_other_agents_observed"]), cfg["skill_levels"],
            bool(cfg["runner_exits_game_after_tagged"]), st["still_in_the_game"],
            bool(cfg["use_full_observation"]), obs, actions, nd, nid,
            st["nearest_neighbor_ids"], rewards, cfg["step_rewards"], st["num_runners"],
            float(cfg["distance_margin_for_reward"]), float(cfg["tag_reward_for_tagger"]),
            float(cfg["tag_penalty_for_runner"]),
            float(cfg["end_of_game_reward_for_runner"]), st["_done_"], st["_timestep_"],
            int(self.N), int(cfg["episode_length"]),
        ]
        types = [_vp] * 7 + [_cf, _cf, _vp, _vp, _cf, _ci, _vp, _cb, _vp, _cb] + \
                [_vp] * 8 + [_cf] * 4 + [_vp, _vp, _ci, _ci]
        self.launch("CudaTagContinuousStep", args, types)

    def tag_gridworld_step(self, loc_x, loc_y, actions, done, rewards, obs, cfg, timestep,
                           episode_length, index_to_action):
        self.set_constant("kIndexToActionArr", np.asarray(index_to_action, np.int32))
        args = [loc_x, loc_y, actions, done, rewards, obs,
                float(cfg["wall_hit_penalty"]), float(cfg["tag_reward_for_tagger"]),
                float(cfg["tag_penalty_for_runner"]), float(cfg["step_cost_for_tagger"]),
                bool(cfg["use_full_observation"]), int(cfg["world_boundary"]), timestep,
                int(episode_length)]
        types = [_vp] * 6 + [_cf] * 4 + [_cb, _ci, _vp, _ci]
        self.launch("CudaTagGridWorldStep", args, types,
                    block=(self.N, 1, 1), grid=(self.E, 1, 1))

    def init_random(self, seed):
        # the reference `new`s one curandState per thread on the device heap
        # (core/random.cu:14-23): make the heap big enough for E*N states
        need = self.E * self.N * 64 + (8 << 20)
        (err,) = cu.cuCtxSetLimit(cu.CUlimit.CU_LIMIT_MALLOC_HEAP_SIZE, need)
        self.launch("init_random", [int(seed)], [_ci])
        self._random_ready = True

    def sample_actions(self, distr, action_indices, cum_distr, n_agents, n_actions,
                       use_argmax=0):
        assert self._random_ready or use_argmax
        block = ((n_agents - 1) // self.bpe + 1, 1, 1)
        self.launch("sample_actions",
                    [distr, action_indices, cum_distr, int(n_agents), int(n_actions),
                     int(use_argmax)], [_vp, _vp, _vp, _ci, _ci, _ci], block=block)

    def reset_when_done(self, data, ref, done, shape, force_reset=0):
        """reset_in_{float,int}_when_done_{2d,3d} chosen like
        pycuda_function_manager.py:686-734."""
        kind = "float" if data.dtype == torch.float32 else "int"
        if len(shape) >= 3:
            agent_dim, feat = int(shape[1]), int(np.prod(shape[2:]))
            self.launch(f"reset_in_{kind}_when_done_3d",
                        [data, ref, done, agent_dim, feat, int(force_reset)],
                        [_vp, _vp, _vp, _ci, _ci, _ci],
                        block=((agent_dim - 1) // self.bpe + 1, 1, 1))
        else:
            feat = int(shape[1]) if len(shape) == 2 else 1
            self.launch(f"reset_in_{kind}_when_done_2d",
                        [data, ref, done, feat, int(force_reset)],
                        [_vp, _vp, _vp, _ci, _ci],
                        block=((feat - 1) // self.bpe + 1, 1, 1))

    def undo_done(self, done, timestep, force_reset=0):
        self.launch("undo_done_flag_and_reset_timestep", [done, timestep, int(force_reset)],
                    [_vp, _vp, _ci], block=(1, 1, 1))
