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
