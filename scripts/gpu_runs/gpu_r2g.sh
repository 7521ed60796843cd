#!/bin/bash
# v2 with alive-first compaction: parity (both variants), A/B bench, ncu
mkdir -p gpurun_out
WDB_OPTIONS=tc_variant=2 timeout 900 python -m pytest tests/test_gpu_envs.py tests/test_gpu_rollout.py -x -q --tb=short 2>&1 | tail -12
WDB_OPTIONS=tc_variant=2,tc_v2_threads=224 timeout 900 python -m pytest tests/test_gpu_envs.py tests/test_gpu_rollout.py -x -q --tb=short 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_envs.py tests/test_gpu_rollout.py tests/test_gpu_core.py tests/test_gpu_mlp.py tests/test_gpu_single_agent_rollout.py -x -q --tb=short 2>&1 | tail -4
for v in "tc_variant=1" "tc_variant=2" "tc_variant=2,tc_v2_threads=224"; do
  WDB_OPTIONS=$v timeout 300 python bench.py --steps 200 --warmup 50 --skip-cpu-baseline --skip-ref-gpu > "gpurun_out/r2g_$v.json" 2> "gpurun_out/r2g_$v.err"
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2g_{v}.json").read().strip().splitlines()[-1])
    print(v, round(d["value"]/1e6, 1), "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "in_rollout", d["roofline"]["kernel_ms_in_rollout"], d["kernel_stats"])
except Exception as e:
    print(v, "failed", e); print(open(f"gpurun_out/r2g_{v}.err").read()[-1500:])
PY
done
WDB_OPTIONS=tc_variant=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_small_v2_kernel -s 60 -c 1 \
  -o gpurun_out/prof_v2_r2g -f python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --skip-ref-gpu > gpurun_out/ncu_v2_r2g.log 2>&1
tail -n 2 gpurun_out/ncu_v2_r2g.log
exit 0
