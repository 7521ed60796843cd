"""Generate tests/golden/classic_control_cudasim.npz from the REAL reference kernels.

Run in the build container only (needs /root/reference and numba; no GPU):

    python tests/golden/make_classic_control_golden.py

The reference's single-agent classic-control steps exist only as numba CUDA kernels
(example_envs/single_agent/classic_control/*/*_step_numba.py).  numba ships a CUDA
*simulator* (NUMBA_ENABLE_CUDASIM=1) that executes such kernels on the CPU, one Python
thread per CUDA thread -- this script imports each kernel module from its file under
/root/reference (nothing is copied), launches the kernel exactly as the reference does
(`kernel[n_envs, 1](...)`, float32/int32 arrays, np.float32/np.int32 scalars) on seeded
inputs that cover the clip / wrap / terminal branches, and records inputs and outputs.

Caveat, stated where the fixture is used (tests/test_oracle_cpu.py): the simulator runs the
Python source with NumPy scalar semantics, not numba's compiled typing -- e.g. `3 *
position` stays float32 under NumPy 2 where compiled numba promotes to float64 -- so the
fixture pins the ALGORITHM (branches, clipping, wrap, reward / done logic, argument order)
to 1e-5, not the last bit.  The last bit is pinned on the GPU box against the compiled
reference kernels (oracle/build_ref_numba.py, tests/test_gpu_classic_control.py).
"""
import importlib.util
import os
import sys

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.build_ref_numba import KERNELS, REF  # noqa: E402

E, STEPS, EP_LEN = 96, 6, 40
PI = np.pi

CONSTS = {
    "cartpole": [9.8, 0.1, 1.1, 0.5, 0.05, 10.0, 0.02, 12 * 2 * np.pi / 360, 2.4],
    "mountain_car": [-1.2, 0.6, 0.07, 0.5, 0.0, 0.001, 0.0025],
    "continuous_mountain_car": [-1.0, 1.0, -1.2, 0.6, 0.07, 0.45, 0.0, 0.0015],
    "pendulum": [],
    "acrobot": [],
}


def _mc_states(rs, n):
    s = np.stack([rs.uniform(-1.25, 0.65, n), rs.uniform(-0.08, 0.08, n)], -1)
    s = np.clip(s, [-1.2, -0.07], [0.6, 0.07])
    s[:6, 0] = -1.2
    s[6:12, 0] = rs.uniform(0.44, 0.6, 6)
    return s


DRAW = {
    "cartpole": (4, 4, np.int32,
                 lambda rs, n: rs.uniform(-1, 1, (n, 4)) * [2.6, 3.0, 0.25, 3.0],
                 lambda rs, n: rs.randint(0, 2, n)),
    "mountain_car": (2, 2, np.int32, _mc_states, lambda rs, n: rs.randint(0, 3, n)),
    "continuous_mountain_car": (2, 2, np.float32, _mc_states,
                                lambda rs, n: rs.uniform(-1.5, 1.5, n)),
    "pendulum": (2, 3, np.float32, lambda rs, n: rs.uniform(-1, 1, (n, 2)) * [12.0, 8.0],
                 lambda rs, n: rs.uniform(-3.0, 3.0, n)),
    "acrobot": (4, 6, np.int32,
                lambda rs, n: rs.uniform(-1, 1, (n, 4)) * [PI, PI, 4 * PI, 9 * PI],
                lambda rs, n: rs.randint(0, 3, n)),
}


def main():
    from numba import cuda

    assert cuda.is_available() and type(cuda).__name__ == "module"
    out = {}
    for name, (rel, symbol, kinds) in KERNELS.items():
        spec = importlib.util.spec_from_file_location(f"wd_ref_sim_{name}", os.path.join(REF, rel))
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        kernel = getattr(module, symbol)
        # helpers such as `_clip` are declared with a bare @cuda.jit; compiled numba turns them
        # into device functions when a kernel calls them, the simulator does not -- unwrap
        # them to the plain Python function they decorate (in memory only)
        for attr, obj in list(vars(module).items()):
            if obj is not kernel and type(obj).__name__ == "FakeCUDAKernel":
                setattr(module, attr, obj.fn)
        sdim, odim, adtype, draw_state, draw_action = DRAW[name]
        consts = [np.float32(c) for c in CONSTS[name]]
        rs = np.random.RandomState(1000 + len(name))
        rec = {k: [] for k in ("state_in", "action", "timestep_in", "state_out", "obs", "reward",
                               "done", "timestep_out")}
        for _ in range(STEPS):
            state = draw_state(rs, E).astype(np.float32).reshape(E, 1, sdim)
            action = draw_action(rs, E).astype(adtype).reshape(E, 1, 1)
            ts = rs.randint(0, EP_LEN, E).astype(np.int32)
            ts[:16] = EP_LEN - 1
            rec["state_in"].append(state.copy())
            rec["action"].append(action.copy())
            rec["timestep_in"].append(ts.copy())
            done = np.zeros(E, np.int32)
            reward = np.full((E, 1), 7.0, np.float32)
            obs = np.full((E, 1, odim), 7.0, np.float32)
            kernel[E, 1](state, action, done, reward, obs, *consts, ts, np.int32(EP_LEN))
            rec["state_out"].append(state.copy())
            rec["obs"].append(obs)
            rec["reward"].append(reward)
            rec["done"].append(done)
            rec["timestep_out"].append(ts.copy())
        for k, v in rec.items():
            out[f"{name}__{k}"] = np.stack(v)
        out[f"{name}__consts"] = np.asarray(CONSTS[name], np.float32)
        print(name, "done values:", np.unique(out[f"{name}__done"]))
    out["episode_length"] = np.int32(EP_LEN)
    path = os.path.join(HERE, "classic_control_cudasim.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
