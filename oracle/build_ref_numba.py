"""Recipe: compile the REFERENCE's own numba kernels into oracle/_ref/numba_*.{ptx,cubin}.

TEST INFRASTRUCTURE.  The single-agent classic-control steps of the reference exist
only as numba kernels (example_envs/single_agent/classic_control/*/*_step_numba.py);
their arithmetic is whatever numba's type inference makes of the Python source
(float32 array loads meeting float64 literals, int64 * float32 -> float64, libdevice
math), which no reading of the source pins down with certainty.  This script lets the
reference's own compiler decide: it imports each kernel module FROM ITS FILE under
/root/reference (nothing is copied), runs `numba.cuda.compile_ptx` (NVVM + libdevice; no
GPU needed) with the argument types the reference's data manager produces
(warp_drive/managers/data_manager.py:107-128, 263-269: float32/int32 arrays, np.float32 /
np.int32 scalars) and assembles the PTX with `ptxas -arch=sm_100a`.  Outputs land in
oracle/_ref/ (git-ignored, shipped to the GPU box) next to a manifest naming each entry
point and its flattened parameter list; oracle/ref_numba.py launches them through the CUDA
driver API, the same way oracle/ref_cuda.py launches the reference's CUDA-C kernels.

    python oracle/build_ref_numba.py
"""
import importlib.util
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
_CC = (10, 0)

_CLASSIC = "example_envs/single_agent/classic_control"

# name -> (source file, kernel symbol, argument kinds).  Kinds: f3/i3/f2/i1 = C-contiguous
# float32/int32 arrays of that rank, f/i = float32/int32 scalar.  Argument order is the
# kernel's own signature (the *_step_numba.py files), cited per env in oracle/ref_numba.py.
KERNELS = {
    "cartpole": (
        f"{_CLASSIC}/cartpole/cartpole_step_numba.py",
        "NumbaClassicControlCartPoleEnvStep",
        ["f3", "i3", "i1", "f2", "f3"] + ["f"] * 9 + ["i1", "i"]),
    "mountain_car": (
        f"{_CLASSIC}/mountain_car/mountain_car_step_numba.py",
        "NumbaClassicControlMountainCarEnvStep",
        ["f3", "i3", "i1", "f2", "f3"] + ["f"] * 7 + ["i1", "i"]),
    "continuous_mountain_car": (
        f"{_CLASSIC}/continuous_mountain_car/continuous_mountain_car_step_numba.py",
        "NumbaClassicControlContinuousMountainCarEnvStep",
        ["f3", "f3", "i1", "f2", "f3"] + ["f"] * 8 + ["i1", "i"]),
    "acrobot": (
        f"{_CLASSIC}/acrobot/acrobot_step_numba.py",
        "NumbaClassicControlAcrobotEnvStep",
        ["f3", "i3", "i1", "f2", "f3", "i1", "i"]),
    "pendulum": (
        f"{_CLASSIC}/pendulum/pendulum_step_numba.py",
        "NumbaClassicControlPendulumEnvStep",
        ["f3", "f3", "i1", "f2", "f3", "i1", "i"]),
}


def _numba_types():
    from numba import float32, int32

    return {
        "f3": float32[:, :, ::1], "i3": int32[:, :, ::1], "f2": float32[:, ::1],
        "i1": int32[::1], "f": float32, "i": int32,
    }


def _no_device_needed():
    """numba asks the current device for its compute capability when a kernel calls
    another @cuda.jit function (numba/cuda/dispatcher.py compile_device); there is no
    driver in the build container, so answer with the target we compile for."""
    import numba.cuda.dispatcher as dispatcher

    class _Target:
        compute_capability = _CC

    dispatcher.get_current_device = lambda: _Target()


def ptx_path(name):
    return os.path.join(OUT, f"numba_{name}.ptx")


def cubin_path(name):
    return os.path.join(OUT, f"numba_{name}.cubin")


def manifest_path():
    return os.path.join(OUT, "numba_manifest.json")


def build_all(force=False):
    if not os.path.isdir(REF):
        raise FileNotFoundError(
            f"{REF} is only present in the build container; the GPU box uses the "
            "prebuilt oracle/_ref/numba_*.cubin that gpurun ships")
    os.makedirs(OUT, exist_ok=True)
    if (not force and os.path.exists(manifest_path())
            and all(os.path.exists(cubin_path(n)) for n in KERNELS)):
        return manifest_path()
    from numba import cuda

    _no_device_needed()
    types = _numba_types()
    manifest = {}
    for name, (rel, symbol, kinds) in KERNELS.items():
        spec = importlib.util.spec_from_file_location(f"wd_ref_{name}", os.path.join(REF, rel))
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        kernel = getattr(module, symbol)
        ptx, _ = cuda.compile_ptx(kernel.py_func, tuple(types[k] for k in kinds), cc=_CC)
        entry = [ln.split("(")[0].split()[-1] for ln in ptx.splitlines()
                 if ln.startswith(".visible .entry")]
        assert len(entry) == 1, entry
        with open(ptx_path(name), "w") as fp:
            fp.write(ptx)
        subprocess.run(["ptxas", "-arch=sm_100a", ptx_path(name), "-o", cubin_path(name)],
                       check=True)
        manifest[name] = {"entry": entry[0], "symbol": symbol, "source": rel, "args": kinds}
    with open(manifest_path(), "w") as fp:
        json.dump(manifest, fp, indent=1)
    return manifest_path()


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
