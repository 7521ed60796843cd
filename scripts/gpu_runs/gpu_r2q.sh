#!/bin/bash
# window-limited network path of the cluster kernel: parity, config-4 bench, fresh ncu capture
mkdir -p gpurun_out
O=gpurun_out/r2q
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_rollout.py tests/test_gpu_envs.py -m gpu -q --tb=short -x 2>&1 | tail -6 | tee ${O}_tests.log
timeout 600 python bench.py --config 4 --steps 8 --warmup 4 --skip-cpu-baseline --skip-ref-gpu > ${O}_c4_b2.json 2> ${O}_c4_b2.err
timeout 600 python bench.py --config 4 --blocks-per-env 1 --steps 8 --warmup 4 --skip-cpu-baseline --skip-ref-gpu > ${O}_c4_b1.json 2> ${O}_c4_b1.err
python - <<'PY'
import json
for f in ("c4_b2", "c4_b1"):
    try:
        d = json.loads(open(f"gpurun_out/r2q_{f}.json").read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "value", round(d["value"] / 1e6, 2), "M/s ms/step", round(d["ms_per_step"], 5),
              "kernel_ms", r.get("kernel_ms"), d.get("kernel_stats"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2q_{f}.err").read()[-1500:])
PY
timeout 400 ncu --set full --import-source on --clock-control none -k regex:tc_wide_kernel -s 6 -c 1 \
  -o ${O}_prof_wide -f python bench.py --config 4 --steps 4 --warmup 3 --skip-cpu-baseline --skip-ref-gpu > ${O}_ncu_wide.log 2>&1
tail -n 2 ${O}_ncu_wide.log
exit 0
