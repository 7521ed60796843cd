#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_update.py -m gpu -q --tb=short 2>&1 | tail -4 | tee ${O}_tests.log
timeout 600 python bench.py --mode train --steps 40 --warmup 10 > ${O}_train.json 2> ${O}_train.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2w_train.json").read().strip().splitlines()[-1])
    t = d.get("train") or {}
    print("train value", round(d["value"]/1e6,2), "M/s iteration_ms", t.get("iteration_ms"), "rollout", t.get("rollout_ms"), "update", t.get("update_ms"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2w_train.err").read()[-2000:])
PY
timeout 600 python scripts/profile_update.py > ${O}_update_profile.txt 2> ${O}_prof_err.txt
head -32 ${O}_update_profile.txt | cut -c1-170
exit 0
