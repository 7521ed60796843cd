"""ctypes binding of libwdb200.so (the C ABI declared in include/wdb200.h).

This is the ONLY compute backend of the package: if the shared library is missing the
import of any kernel-facing module fails loudly -- there is no Python / CPU fallback.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwdb200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_ll = ctypes.c_longlong
_ull = ctypes.c_ulonglong


class ResetDesc(ctypes.Structure):
    """struct wdb_reset_desc (include/wdb200.h)."""

    _fields_ = [("dst", _vp), ("ref", _vp), ("bytes_per_env", _ll), ("pool_rows", _ll)]


_SIGNATURES = {
    "wdb_abi_version": (_i, []),
    "wdb_error_string": (ctypes.c_char_p, [_i]),
    "wdb_launch_count": (_ll, []),
    "wdb_rng_state_bytes": (_ll, [_ll]),
    "wdb_rng_init": (_i, [_vp, _vp, _ll, _ull]),
    "wdb_rng_draw_u32x4": (_i, [_vp, _vp, _vp, _ll]),
    "wdb_sample_actions": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "wdb_sample_ou_process": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _vp]),
    "wdb_reset_when_done": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "wdb_reset_log_mask": (_i, [_vp, _vp, _i]),
    "wdb_update_log_mask": (_i, [_vp, _vp, _i, _i]),
    "wdb_log_one_step": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i]),
    "wdb_testkernel": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _f, _i, _i, _i]),
    "wdb_tag_gridworld_step": (
        _i,
        [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _i, _vp, _i, _vp],
    ),
    "wdb_tag_continuous_step": (
        _i,
        [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _f, _i,
         _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _vp,
         _vp, _i, _vp],
    ),
    "wdb_cartpole_step": (
        _i,
        [_vp, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _f, _f, _f, _vp, _i],
    ),
}

_lib = None


def exported_symbols():
    """Names every build must export (== the functions include/wdb200.h declares)."""
    return sorted(_SIGNATURES)


def load():
    """dlopen libwdb200.so and type its entry points.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m warp_drive_b200.build` "
            "(or __graft_entry__.build()).  warp_drive_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the build lost a symbol
        fn.restype = restype
        fn.argtypes = argtypes
    assert lib.wdb_abi_version() == 1
    _lib = lib
    return lib


def check(err, what=""):
    if err != 0:
        msg = load().wdb_error_string(err).decode()
        raise RuntimeError(f"libwdb200 {what} failed: cudaError {err} ({msg})")


def ptr(t):
    """Device pointer of a torch CUDA tensor / DeviceArray handle (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "tensor"):
        t = t.tensor
    if torch.is_tensor(t):
        if not t.is_cuda:
            raise RuntimeError(
                "libwdb200 kernels need CUDA tensors; got a CPU tensor "
                "(warp_drive_b200 has no CPU fallback on the kernel path)"
            )
        if not t.is_contiguous():
            raise RuntimeError("libwdb200 kernels need C-contiguous tensors")
        return t.data_ptr()
    if hasattr(t, "gpudata"):  # pycuda-style pointer holder
        return int(t.gpudata)
    if hasattr(t, "data_ptr"):
        return int(t.data_ptr())
    if isinstance(t, int):
        return t
    raise TypeError(f"cannot take a device pointer of {type(t)}")


def stream_ptr(stream=None):
    """cudaStream_t of torch's current stream, so launches order with torch work."""
    if stream is None:
        stream = torch.cuda.current_stream()
    return stream.cuda_stream


def launch_count():
    return int(load().wdb_launch_count())
