#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2y
timeout 900 python -m pytest tests/test_gpu_single_agent_rollout.py tests/test_gpu_update.py tests/test_gpu_training.py -m gpu -q --tb=short 2>&1 | tail -4 | tee ${O}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke ok" | tee ${O}_smoke.log
timeout 600 python bench.py --config 3 > ${O}_c3.json 2> ${O}_c3.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2y_c3.json").read().strip().splitlines()[-1])
    print("c3 value", round(d["value"]/1e6,2), "M/s ms/step", d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2y_c3.err").read()[-2000:])
PY
exit 0
