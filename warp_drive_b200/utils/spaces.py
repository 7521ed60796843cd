"""Minimal observation/action space types.

The reference depends on `gym.spaces` (not installed in this image, SURVEY.md section
8c).  The rollout path only needs the *shape bookkeeping* of Box / Discrete /
MultiDiscrete / Dict, so the package carries its own small types and uses gym's or
gymnasium's classes instead when one of them is importable.
"""
import numpy as np

try:  # pragma: no cover - not available in the build image
    from gymnasium.spaces import Box, Dict, Discrete, MultiDiscrete, Space  # noqa: F401
except ImportError:  # pragma: no cover
    try:
        from gym.spaces import Box, Dict, Discrete, MultiDiscrete, Space  # noqa: F401
    except ImportError:

        class Space:
            shape = None
            dtype = None

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.dtype = np.dtype(dtype)
                if shape is None:
                    shape = np.shape(low)
                self.shape = tuple(int(s) for s in shape)
                self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
                self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)

            def __repr__(self):
                return f"Box{self.shape}"

        class Discrete(Space):
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()
                self.dtype = np.dtype(np.int64)

            def __repr__(self):
                return f"Discrete({self.n})"

        class MultiDiscrete(Space):
            def __init__(self, nvec):
                self.nvec = np.asarray(nvec, dtype=np.int64)
                self.shape = self.nvec.shape
                self.dtype = np.dtype(np.int64)

            def __repr__(self):
                return f"MultiDiscrete({self.nvec.tolist()})"

        class Dict(Space, dict):
            def __init__(self, spaces=None):
                dict.__init__(self, spaces or {})

            @property
            def spaces(self):
                return self


def obs_dict_to_spaces(obs):
    """{agent_id: array | nested dict} -> Dict of Box spaces.
    Same contract as warp_drive/utils/recursive_obs_dict_to_spaces_dict.py:13."""
    assert isinstance(obs, dict)
    out = {}
    for key, val in obs.items():
        if isinstance(val, dict):
            out[key] = obs_dict_to_spaces(val)
            continue
        arr = np.asarray([val]) if np.isscalar(val) else np.asarray(val)
        if arr.dtype.kind not in "fiub":
            raise TypeError(f"unsupported observation type for key {key}: {arr.dtype}")
        if arr.dtype.kind == "f":
            bound = float(np.finfo(arr.dtype).max) / 2
        else:
            bound = float(np.iinfo(arr.dtype).max // 2) if arr.dtype.kind in "iu" else 1.0
        out[key] = Box(low=-bound, high=bound, shape=arr.shape, dtype=arr.dtype)
    return Dict(out)


# reference name
recursive_obs_dict_to_spaces_dict = obs_dict_to_spaces
