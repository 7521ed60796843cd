#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_single_agent_rollout.py tests/test_gpu_envs.py -x -q --tb=short 2>&1 | tail -6
for B in 2 1 4; do
  timeout 600 python bench.py --config 4 --blocks-per-env $B --steps 8 --warmup 4 --skip-cpu-baseline --skip-ref-gpu > gpurun_out/r2k_c4_b$B.json 2> gpurun_out/r2k_c4_b$B.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2k_c4_b$B.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("C4 bpe=$B value", round(d["value"]/1e6,1), "M/s ms/step", round(d["ms_per_step"],4), "kernel_ms", round(r["kernel_ms"],4), "frac", round(r["frac"],4))
except Exception as e:
    print("C4 bpe=$B failed", e); print(open("gpurun_out/r2k_c4_b$B.err").read()[-2000:])
PY
done
timeout 600 python bench.py --config 3 --skip-cpu-baseline > gpurun_out/r2k_c3.json 2> gpurun_out/r2k_c3.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2k_c3.json").read().strip().splitlines()[-1])
    print("C3 value", round(d["value"]/1e6,1), "M/s ms/step", round(d["ms_per_step"],5), "us/timestep", round(d["roofline"]["kernel_us_per_timestep"],2), "e2e", round(d["e2e"]["value"]/1e6,1))
except Exception as e:
    print("C3 failed", e); print(open("gpurun_out/r2k_c3.err").read()[-2000:])
PY
exit 0
