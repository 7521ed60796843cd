#!/bin/bash
# round-2 first GPU pass: full gpu test-suite, driver-config bench (x2, for repeatability),
# default bench with the reference-GPU anchor.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short --durations=6 2>&1 | tail -25
for i in 1 2; do
  timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu-baseline --skip-ref-gpu > gpurun_out/r2a_drv_$i.json 2> gpurun_out/r2a_drv_$i.err
done
timeout 600 python bench.py > gpurun_out/r2a_default.json 2> gpurun_out/r2a_default.err
python - <<'PY'
import json
for f in ("r2a_drv_1", "r2a_drv_2", "r2a_default"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]/1e6, 1), "ms/step", round(d["ms_per_step"], 4), "reps", [round(x, 4) for x in d["ms_per_step_reps"]],
              "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"]/1e6, 1))
        if "ref_gpu" in d: print("  ref_gpu", json.dumps(d["ref_gpu"]))
        if d["roofline"].get("issue"): print("  issue", d["roofline"]["issue"])
    except Exception as e:
        print(f, "parse failed", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
exit 0
