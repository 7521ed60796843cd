"""CUDADataManager: the named device-array registry of the rollout engine.

Same public contract as the reference's CUDADataManager / PyCUDADataManager
(warp_drive/managers/data_manager.py:17-485, pycuda_managers/pycuda_data_manager.py),
re-designed for one data store: EVERY array is a torch CUDA tensor resident in HBM
(the reference mixes raw pycuda allocations with torch tensors), so
`data_on_device_via_torch(name)` works for every array and kernels receive
`tensor.data_ptr()` through the C ABI.

Kept semantics (they are what env classes and the trainer rely on):
  * 64-bit host data is down-cast to 32-bit at push (data_manager.py:263-269)
  * `save_copy_and_apply_at_reset` registers `{name}_at_reset` (+ reset list order)
  * `log_data_across_episode` registers `{name}_for_log` shaped [episode_length+1, ...]
  * reset pools (`is_reset_pool`, `reset_target`), scalars returned as np.int32/float32
  * reserved arrays `_done_` (torch accessible), `_timestep_`, `_log_mask_`
"""
import logging
from typing import Dict, Optional

import numpy as np
import torch

from warp_drive_b200.utils.data_feed import DataFeed


class DeviceArray:
    """Handle of one named device array (what `device_data(name)` returns).
    Quacks like the reference's pointer holders: `.gpudata` / `.data_ptr()`."""

    __slots__ = ("name", "tensor")

    def __init__(self, name, tensor):
        self.name = name
        self.tensor = tensor

    @property
    def gpudata(self):
        return self.tensor.data_ptr()

    def data_ptr(self):
        return self.tensor.data_ptr()

    def __repr__(self):
        return f"DeviceArray({self.name}, {tuple(self.tensor.shape)}, {self.tensor.dtype})"


def _default_device():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "CUDADataManager needs a CUDA device (pass device='cpu' explicitly only to "
            "unit-test the registry logic; kernels never run on CPU tensors)"
        )
    return torch.device("cuda", torch.cuda.current_device())


class CUDADataManager:
    def __init__(self, num_agents: int = None, num_envs: int = None,
                 blocks_per_env: int = 1, episode_length: int = None, device=None):
        assert num_agents is not None and num_envs is not None
        assert blocks_per_env is not None and episode_length is not None
        self.device = torch.device(device) if device is not None else _default_device()
        self._meta_info = {}
        self._host_data = {}
        self._device_data_pointer = {}
        self._device_data_via_torch = {}
        self._torch_accessible = set()
        self._scalar_data_list = []
        self._reset_data_list = []
        self._reset_target_to_pool = {}
        self._log_data_list = []
        self._shared_constants = {}
        self._shape = {}
        self._dtype = {}
        self.add_meta_info({
            "n_agents": num_agents, "episode_length": episode_length,
            "n_envs": num_envs, "blocks_per_env": blocks_per_env,
        })
        # reserved arrays (data_manager.py:75-105)
        feed = DataFeed()
        feed.add_data(name="_log_mask_", data=np.zeros(episode_length + 1, np.int32))
        self.push_data_to_device(feed)
        feed = DataFeed()
        feed.add_data(name="_done_", data=np.zeros(num_envs, np.int32))
        self.push_data_to_device(feed, torch_accessible=True)
        feed = DataFeed()
        feed.add_data(name="_timestep_", data=np.zeros(num_envs, np.int32))
        self.push_data_to_device(feed)

    # ------------------------------------------------------------------ registration
    @staticmethod
    def _as_32bit_array(key, value):
        if isinstance(value, np.ndarray):
            array = value if value.flags.c_contiguous else np.ascontiguousarray(value)
        elif isinstance(value, list):
            array = np.array(value, order="C")
        else:
            raise ValueError(f"the data '{key}' needs to be cast to a list or an array")
        if array.dtype == np.float64:
            logging.warning(f"CUDADataManager casts the data '{key}' from float64 to float32")
            array = array.astype(np.float32)
        elif array.dtype == np.int64:
            logging.warning(f"CUDADataManager casts the data '{key}' from int64 to int32")
            array = array.astype(np.int32)
        elif array.dtype == np.bool_:
            array = array.astype(np.int32)
        return array

    @staticmethod
    def _as_32bit_scalar(value):
        if isinstance(value, (bool, np.bool_, int, np.integer)):
            return np.int32(value)
        return np.float32(value)

    def add_meta_info(self, meta: Dict):
        assert isinstance(meta, dict)
        for key, value in meta.items():
            assert key not in self._meta_info, f"meta info {key} is already registered"
            assert isinstance(value, (int, np.integer, float, np.floating)), (
                "the meta info only accepts scalar int or float values"
            )
            self._meta_info[key] = self._as_32bit_scalar(value)

    def add_shared_constants(self, constants: Dict):
        for key, value in constants.items():
            assert key not in self._shared_constants, f"shared constant {key} already added"
            if isinstance(value, (np.ndarray, list)):
                array = self._as_32bit_array(key, value)
                self._shared_constants[key] = array
                self._shape[key] = array.shape
                self._dtype[key] = array.dtype.name
            elif isinstance(value, (int, np.integer, float, np.floating)):
                self._shared_constants[key] = self._as_32bit_scalar(value)
                self._shape[key] = ()
                self._dtype[key] = self._shared_constants[key].dtype.name
            else:
                raise ValueError(f"shared constant '{key}' must be a scalar, list or array")

    def push_data_to_device(self, data: Dict, torch_accessible: bool = False):
        assert isinstance(data, dict)
        for key, content in data.items():
            assert key not in self._host_data, f"data {key} is already registered"
            value = content["data"]
            attrs = content["attributes"]
            is_pool = bool(attrs.get("is_reset_pool", False))
            save_copy = bool(attrs.get("save_copy_and_apply_at_reset", False)) and not is_pool
            log_it = bool(attrs.get("log_data_across_episode", False)) and not is_pool

            if isinstance(value, (np.ndarray, list)):
                assert key not in self._device_data_pointer
                if is_pool:
                    target = attrs["reset_target"]
                    assert target not in self._reset_target_to_pool
                    assert target not in self._reset_data_list
                    self._reset_target_to_pool[target] = key
                array = self._as_32bit_array(key, value)
                self._host_data[key] = array
                self._shape[key] = array.shape
                self._dtype[key] = array.dtype.name
                self._to_device(key, torch_accessible=torch_accessible)
                if save_copy:
                    assert key not in self._reset_data_list
                    assert key not in self._reset_target_to_pool
                    at_reset = f"{key}_at_reset"
                    self._shape[at_reset] = array.shape
                    self._dtype[at_reset] = array.dtype.name
                    self._to_device(key, name_on_device=at_reset)
                    self._reset_data_list.append(key)
                if log_it:
                    assert key not in self._log_data_list
                    assert array.shape[0] == self.meta_info("n_envs")
                    assert array.shape[1] == self.meta_info("n_agents")
                    for_log = f"{key}_for_log"
                    self._host_data[for_log] = np.zeros(
                        (int(self.meta_info("episode_length")) + 1, *array.shape[1:]),
                        dtype=array.dtype,
                    )
                    self._shape[for_log] = self._host_data[for_log].shape
                    self._dtype[for_log] = array.dtype.name
                    self._to_device(for_log)
                    self._log_data_list.append(key)
            elif isinstance(value, (bool, np.bool_, int, np.integer, float, np.floating)):
                assert key not in self._scalar_data_list
                self._host_data[key] = self._as_32bit_scalar(value)
                self._shape[key] = ()
                self._dtype[key] = self._host_data[key].dtype.name
                self._scalar_data_list.append(key)
            else:
                raise ValueError(f"the data '{key}' must be a scalar, list or array")

    def _to_device(self, name, name_on_device: Optional[str] = None,
                   torch_accessible: bool = False):
        host = self._host_data[name]
        dev_name = name_on_device or name
        assert dev_name not in self._device_data_pointer
        tensor = torch.from_numpy(np.ascontiguousarray(host)).to(self.device).contiguous()
        if tensor.device == torch.device("cpu"):
            tensor = tensor.clone()  # never alias the host copy
        self._device_data_via_torch[dev_name] = tensor
        self._device_data_pointer[dev_name] = DeviceArray(dev_name, tensor)
        if torch_accessible:
            self._torch_accessible.add(dev_name)

    # ------------------------------------------------------------------ access
    def pull_data_from_device(self, name: str):
        if name in self._scalar_data_list:
            return self._host_data[name]
        assert name in self._device_data_via_torch, f"{name} is not on the device"
        return self._device_data_via_torch[name].detach().cpu().numpy()

    def data_on_device_via_torch(self, name: str) -> torch.Tensor:
        assert name in self._device_data_via_torch, f"{name} is not on the device"
        return self._device_data_via_torch[name]

    def reset_device(self, name: Optional[str] = None):
        """Host -> device copy of the values registered at push time."""
        names = [name] if name is not None else [
            k for k in self._host_data if k in self._device_data_via_torch
        ]
        for key in names:
            assert key in self._device_data_via_torch and key in self._host_data
            self._device_data_via_torch[key].copy_(torch.from_numpy(self._host_data[key]))

    def meta_info(self, name: str):
        assert name in self._meta_info
        return self._meta_info[name]

    def shared_constant(self, name: str):
        assert name in self._shared_constants
        return self._shared_constants[name]

    def device_data(self, name: str):
        if name in self._scalar_data_list:
            return self._host_data[name]
        assert name in self._device_data_pointer, f"{name} is not on the device"
        return self._device_data_pointer[name]

    def is_data_on_device(self, name: str) -> bool:
        return name in self._device_data_pointer

    def is_data_on_device_via_torch(self, name: str) -> bool:
        # every array is a torch tensor here; the reference's distinction is kept only
        # for arrays pushed with torch_accessible=True so reference tests read the same
        return name in self._device_data_pointer and name in self._device_data_via_torch

    def get_shape(self, name: str):
        assert name in self._shape
        return self._shape[name]

    def get_dtype(self, name: str):
        assert name in self._dtype
        return self._dtype[name]

    def get_reset_pool(self, name: str):
        assert name in self._reset_target_to_pool
        return self._reset_target_to_pool[name]

    @property
    def host_data(self):
        return self._host_data

    @property
    def scalar_data_list(self):
        return self._scalar_data_list

    @property
    def reset_data_list(self):
        return self._reset_data_list

    @property
    def reset_target_to_pool(self):
        return self._reset_target_to_pool

    @property
    def log_data_list(self):
        return self._log_data_list

    @property
    def device_data_via_torch(self):
        return self._device_data_via_torch


# names a reference user imports
PyCUDADataManager = CUDADataManager
NumbaDataManager = CUDADataManager
B200DataManager = CUDADataManager
