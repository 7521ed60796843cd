#!/bin/bash
mkdir -p gpurun_out
WDB_OPTIONS=tc_variant=2,tc_v2_threads=128 timeout 900 python -m pytest tests/test_gpu_envs.py tests/test_gpu_rollout.py -x -q --tb=short 2>&1 | tail -4
for v in "tc_variant=1" "tc_variant=2,tc_v2_threads=128" "tc_variant=2"; do
  WDB_OPTIONS=$v timeout 300 python bench.py --steps 200 --warmup 50 --skip-cpu-baseline --skip-ref-gpu > "gpurun_out/r2h_$v.json" 2> "gpurun_out/r2h_$v.err"
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2h_{v}.json").read().strip().splitlines()[-1])
    print(v, round(d["value"]/1e6, 1), "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "in_rollout", d["roofline"]["kernel_ms_in_rollout"])
except Exception as e:
    print(v, "failed", e); print(open(f"gpurun_out/r2h_{v}.err").read()[-1500:])
PY
done
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_training.py tests/test_gpu_update.py tests/test_gpu_envs.py -x -q --tb=short -k "not tag_continuous" 2>&1 | tail -6
timeout 600 python bench.py --mode train --steps 40 --warmup 10 > gpurun_out/r2h_train1.json 2> gpurun_out/r2h_train1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2h_train1.json").read().strip().splitlines()[-1])
    print("train1", round(d["value"]/1e6, 1), "M/s", json.dumps(d["train"])[:500])
except Exception as e:
    print("train failed", e); print(open("gpurun_out/r2h_train1.err").read()[-2500:])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
exit 0
