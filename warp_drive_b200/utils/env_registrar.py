"""EnvironmentRegistrar: name -> env class lookup for EnvWrapper(env_name=...).
API of warp_drive/utils/env_registrar.py:4-132.  The reference also records a path to a
user .cu file that it JIT-compiles; with a prebuilt libwdb200.so custom device code is
registered through CUDAFunctionManager.register_function instead."""


class EnvironmentRegistrar:
    _BACKENDS = ("cpu", "pycuda", "numba", "b200")

    def __init__(self):
        self._cpu_envs = {}
        self._device_envs = {}
        self._customized_env_path = {}

    def add(self, env_backend="cpu", cuda_env_src_path=None):
        if isinstance(env_backend, bool):  # legacy use_cuda flag
            env_backend = "pycuda" if env_backend else "cpu"
        assert env_backend in self._BACKENDS, f"unknown env_backend {env_backend}"

        def register(cls):
            name = cls.name.lower()
            table = self._cpu_envs if env_backend == "cpu" else self._device_envs
            if name in table:
                raise KeyError(f"environment {cls.name} is already registered")
            table[name] = cls
            if cuda_env_src_path is not None:
                self._customized_env_path[name] = cuda_env_src_path
            return cls

        return register

    def get(self, name, env_backend="cpu"):
        if isinstance(env_backend, bool):
            env_backend = "pycuda" if env_backend else "cpu"
        table = self._cpu_envs if env_backend == "cpu" else self._device_envs
        key = name.lower()
        if key not in table:
            raise KeyError(f"no environment '{name}' registered for backend {env_backend}")
        return table[key]

    def has_env(self, name, env_backend="cpu"):
        table = self._cpu_envs if env_backend == "cpu" else self._device_envs
        return name.lower() in table

    def get_cuda_env_src_path(self, name, env_backend="pycuda"):
        return self._customized_env_path.get(name.lower())
