#!/bin/bash
# config 4 bench: cluster sizes 2/4/8, window on/off
mkdir -p gpurun_out
for B in 4 2 8; do
  timeout 600 python bench.py --config 4 --blocks-per-env $B --steps 8 --warmup 4 --skip-cpu-baseline $( [ $B != 4 ] && echo --skip-ref-gpu ) > gpurun_out/r2b_c4_b$B.json 2> gpurun_out/r2b_c4_b$B.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2b_c4_b$B.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("C4 bpe=$B value", round(d["value"]/1e6,1), "M/s ms/step", round(d["ms_per_step"],4), "kernel_ms", round(r["kernel_ms"],4), "frac", round(r["frac"],4), "pairs/s", r["pair_evals"]["nominal_pairs_per_s"], "e2e", round(d["e2e"]["value"]/1e6,1), d["kernel_stats"])
    if "ref_gpu" in d: print("  ref_gpu", json.dumps(d["ref_gpu"]))
except Exception as e:
    print("C4 bpe=$B failed", e); print(open("gpurun_out/r2b_c4_b$B.err").read()[-2500:])
PY
done
exit 0
