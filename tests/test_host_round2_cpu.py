"""CPU tests of the host logic added in round 2: NUMA binding of ranks, the bench's shared
`config` object, the wide-kernel / whole-rollout eligibility rules (no GPU calls)."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_TOPO = ("\t\x1b[4mGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\n"
         "GPU0\t X \tNV18\tPIX\t0-31,64-95\t0\t\tN/A\n"
         "GPU1\tNV18\t X \tSYS\t32-63,96-127\t1\t\tN/A\n"
         "NIC0\tPIX\tSYS\t X \t\t\t\t\n\nLegend:\n")


def test_gpu_cpu_affinity_parses_nvidia_smi_topology(monkeypatch):
    from warp_drive_b200.utils import numa

    monkeypatch.setattr(numa.subprocess, "run",
                        lambda *a, **k: types.SimpleNamespace(stdout=_TOPO, returncode=0))
    assert numa.gpu_cpu_affinity(0) == set(range(0, 32)) | set(range(64, 96))
    assert numa.gpu_cpu_affinity(1) == set(range(32, 64)) | set(range(96, 128))
    assert numa.gpu_cpu_affinity(5) is None
    # binding intersects with what the process may use and never raises
    monkeypatch.setattr(numa.os, "sched_getaffinity", lambda pid: {0, 1, 40}, raising=False)
    bound = {}
    monkeypatch.setattr(numa.os, "sched_setaffinity", lambda pid, s: bound.update(s=set(s)),
                        raising=False)
    assert numa.bind_process_to_gpu(0) == [0, 1] and bound["s"] == {0, 1}
    assert numa.bind_process_to_gpu(1) == [40]
    monkeypatch.setattr(numa.subprocess, "run",
                        lambda *a, **k: (_ for _ in ()).throw(FileNotFoundError()))
    assert numa.bind_process_to_gpu(0) is None


def test_bench_arms_print_the_same_config_object():
    """The reference arm (CPU port) and the b200 arm describe the workload with the identical
    `config` object, for every bench config (the driver compares them)."""
    sys.path.insert(0, ROOT)
    import bench

    for cfg_id in (2, 4):
        bench._ACTIVE["config"] = cfg_id
        bench._ACTIVE["blocks_per_env"] = None
        a = bench.workload_config(2000)
        b = bench.workload_config(2000, a["agents"])
        assert a == b and "workload" in a
    bench._ACTIVE["config"] = 2
    assert bench.config3_line_config(10000)["agents"] == 1
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--config", "3", "--steps", "5", "--warmup", "3", "--envs3", "256"],
                         capture_output=True, text=True, timeout=300)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["warmup"] == 3 and line["steps"] == 5
    assert line["config"] == bench.config3_line_config(256)
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0


def test_wide_kernel_shared_memory_budget_formula():
    """The [E, N, N-1] reference scratch is only requested for the packed single-CTA kernel:
    blocks_per_env > 1 and N > 320 (x-binned kernel) keep everything on chip."""
    from warp_drive_b200.envs.tag_continuous import TagContinuous

    kw = dict(num_taggers=24, num_runners=1000, grid_length=64.0, episode_length=10,
              num_other_agents_observed=10, use_full_observation=False, seed=1)
    env = TagContinuous(**kw)
    env.reset()
    env.get_data_dictionary()
    assert not env.allocate_reference_scratch          # N = 1024, K = 10: wide kernel
    env2 = TagContinuous(**dict(kw, num_other_agents_observed=20))
    env2.reset()
    env2.get_data_dictionary()
    assert env2.allocate_reference_scratch              # K + 2 > 16: the packed kernel's exact path


def test_weight_gradient_chunking_equals_plain_product():
    """fused_mlp_train._wgrad: a^T b as a batched GEMM over row chunks + a sum of the partials
    (the layout that streams at the HBM peak on the GPU) -- same result as the plain product,
    and the plain product when the row count does not split."""
    import torch

    from warp_drive_b200.training.models.fused_mlp_train import _wgrad

    g = torch.Generator().manual_seed(0)
    for rows in (16 * 8192, 2 * 8192 + 2, 1000):
        a = torch.randn(rows, 12, generator=g, dtype=torch.float64)
        b = torch.randn(rows, 7, generator=g, dtype=torch.float64)
        assert torch.allclose(_wgrad(a, b), a.t() @ b, rtol=1e-12, atol=1e-9), rows


def test_fused_train_forward_is_not_taken_off_the_gpu():
    """The fused training node needs CUDA float32 tensors: on CPU tensors the model keeps the
    plain module path (and still differentiates)."""
    import torch

    from warp_drive_b200.training.models import fused_mlp_train

    class M(torch.nn.Module):
        is_deterministic = False
        action_mask = None
        output_dims = [3]

        def __init__(self):
            super().__init__()
            self.fc = torch.nn.ModuleDict({
                "0": torch.nn.Sequential(torch.nn.Linear(5, 8), torch.nn.ReLU()),
                "1": torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.ReLU())})
            self.policy_head = torch.nn.ModuleList([torch.nn.Linear(8, 3)])
            self.vf_head = torch.nn.Linear(8, 1)

    assert not fused_mlp_train.supported(M(), torch.randn(4, 5))
