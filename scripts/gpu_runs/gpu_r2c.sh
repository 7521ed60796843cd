#!/bin/bash
# ncu captures of the small-env fused kernel, variant 1 vs 2; wide kernel tests after the smem re-layout
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_envs.py tests/test_gpu_rollout.py -x -q --tb=short 2>&1 | tail -15
WDB_OPTIONS=tc_variant=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:tag_continuous_kernel -s 30 -c 1 \
  -o gpurun_out/prof_v1_r2c -f python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-ref-gpu > gpurun_out/ncu_v1_r2c.log 2>&1
WDB_OPTIONS=tc_variant=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_small_v2_kernel -s 30 -c 1 \
  -o gpurun_out/prof_v2_r2c -f python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-ref-gpu > gpurun_out/ncu_v2_r2c.log 2>&1
tail -n 2 gpurun_out/ncu_v1_r2c.log gpurun_out/ncu_v2_r2c.log
for B in 1 2 4; do
  timeout 600 python bench.py --config 4 --blocks-per-env $B --steps 8 --warmup 4 --skip-cpu-baseline --skip-ref-gpu > gpurun_out/r2c_c4_b$B.json 2> gpurun_out/r2c_c4_b$B.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c_c4_b$B.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("C4 bpe=$B value", round(d["value"]/1e6,1), "M/s ms/step", round(d["ms_per_step"],4), "kernel_ms", round(r["kernel_ms"],4), "frac", round(r["frac"],4))
except Exception as e:
    print("C4 bpe=$B failed", e); print(open("gpurun_out/r2c_c4_b$B.err").read()[-2000:])
PY
done
exit 0
