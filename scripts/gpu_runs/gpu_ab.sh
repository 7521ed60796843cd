#!/bin/bash
# A/B bench runs: bash scripts/gpu_ab.sh <tag> "<bench args 1>" "<bench args 2>" ...
TAG=$1; shift
i=0
for ARGS in "$@"; do
  i=$((i+1))
  timeout 300 python bench.py --steps 100 --warmup 20 --skip-cpu-baseline $ARGS > gpurun_out/ab_${TAG}_$i.json 2> gpurun_out/ab_${TAG}_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${TAG}_$i.json").read().strip().splitlines()[-1])
    print("AB[$ARGS]", round(d["value"]/1e6,1), "M/s  step_ms", round(d["ms_per_step"],4), " kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print("AB[$ARGS] failed", e); print(open("gpurun_out/ab_${TAG}_$i.err").read()[-1500:])
PY
done
