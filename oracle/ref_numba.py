"""GPU-side oracle for the single-agent classic-control steps: the REFERENCE's own numba
kernels, compiled by numba itself (oracle/build_ref_numba.py -> oracle/_ref/numba_*.cubin)
and launched here through the CUDA driver API on torch tensors.

TEST INFRASTRUCTURE.  numba is not needed at run time (the GPU box only loads the prebuilt
cubins); what has to be reproduced is numba's kernel ABI: every array argument is passed
as the flattened struct
    (meminfo*, parent*, nitems:i64, itemsize:i64, data*, shape[ndim]:i64, strides[ndim]:i64)
(numba/core/datamodel/models.py ArrayModel + numba/cuda/dispatcher.py _prepare_args), scalars
by value.  Launch geometry is the reference's: grid = n_envs blocks of 1 thread
(warp_drive/managers/function_manager.py:65-67 with one agent per env).
"""
import ctypes
import json
import os

import torch

try:
    from cuda.bindings import driver as cu
except ImportError:  # older cuda-python
    from cuda import cuda as cu

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

_DTYPES = {"f": torch.float32, "i": torch.int32}


def _ok(res):
    if int(res[0]) != 0:
        raise RuntimeError(f"CUDA driver error {res[0]}")
    return res[1] if len(res) == 2 else res[1:]


def manifest():
    path = os.path.join(_REF, "numba_manifest.json")
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} missing: run `python oracle/build_ref_numba.py` in the build container "
            "(needs /root/reference and numba)")
    with open(path) as fp:
        return json.load(fp)


def available(name):
    return (os.path.exists(os.path.join(_REF, "numba_manifest.json"))
            and os.path.exists(os.path.join(_REF, f"numba_{name}.cubin")))


def _flatten_array(t, kind):
    """numba's flattened array struct for a C-contiguous torch CUDA tensor."""
    ndim = int(kind[1])
    assert t.is_cuda and t.is_contiguous() and t.dim() == ndim, (kind, tuple(t.shape))
    assert t.dtype == _DTYPES[kind[0]], (kind, t.dtype)
    item = t.element_size()
    vals = [0, 0, t.numel(), item, t.data_ptr()] + list(t.shape) + [s * item for s in t.stride()]
    types = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
             ctypes.c_void_p] + [ctypes.c_int64] * (2 * ndim)
    return vals, types


class RefNumbaKernel:
    """One reference numba kernel (`name` = a key of build_ref_numba.KERNELS)."""

    def __init__(self, name):
        info = manifest()[name]
        torch.zeros(1, device="cuda")  # torch's primary context becomes current
        with open(os.path.join(_REF, f"numba_{name}.cubin"), "rb") as fp:
            self._image = fp.read()
        self.module = _ok(cu.cuModuleLoadData(self._image))
        self.function = _ok(cu.cuModuleGetFunction(self.module, info["entry"].encode()))
        self.kinds = info["args"]
        self.symbol = info["symbol"]

    def __call__(self, n_envs, *args):
        assert len(args) == len(self.kinds), (len(args), len(self.kinds))
        vals, types = [], []
        for a, kind in zip(args, self.kinds):
            if kind == "f":
                vals.append(float(a)); types.append(ctypes.c_float)
            elif kind == "i":
                vals.append(int(a)); types.append(ctypes.c_int)
            else:
                v, t = _flatten_array(a, kind)
                vals += v; types += t
        # explicit void*[] of pointers to the argument values (cuLaunchKernel's native form)
        holders = [t(v) for v, t in zip(vals, types)]
        params = (ctypes.c_void_p * len(holders))(*[ctypes.addressof(h) for h in holders])
        stream = torch.cuda.current_stream().cuda_stream
        (err,) = cu.cuLaunchKernel(self.function, int(n_envs), 1, 1, 1, 1, 1, 0, stream,
                                   ctypes.addressof(params), 0)
        if int(err) != 0:
            raise RuntimeError(f"cuLaunchKernel({self.symbol}) failed: {err}")
