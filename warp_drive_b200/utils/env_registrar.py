"""EnvironmentRegistrar: name -> env class lookup for EnvWrapper(env_name=...).
API of warp_drive/utils/env_registrar.py:4-132, including the path of a user `.cu` file:
the reference JIT-compiles it with nvcc + pycuda; here CUDAFunctionManager.
compile_and_load_cuda compiles it for sm_100a (utils/custom_kernels.py) and serves its
kernels next to the prebuilt libwdb200 ones."""


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

    def add_cuda_env_src_path(self, name, cuda_env_src_path, env_backend="pycuda"):
        """Register the ABSOLUTE path of a custom env's CUDA source (reference :88-119)."""
        assert str(cuda_env_src_path).endswith(".cu"), (
            "the customized environment is expected to be a CUDA source code (*.cu)")
        self._customized_env_path[name.lower()] = str(cuda_env_src_path)

    def get_cuda_env_src_path(self, name, env_backend="pycuda"):
        return self._customized_env_path.get(name.lower())
