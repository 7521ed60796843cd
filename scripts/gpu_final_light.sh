#!/bin/bash
# Light closing pass: full GPU suite, smoke, the driver's bench configuration, train mode.
mkdir -p gpurun_out
O=gpurun_out/final3_r2
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 | tee ${O}_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke ok" | tee ${O}_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_driver.json 2> ${O}_bench_driver.err
timeout 600 python bench.py --mode train --steps 40 --warmup 10 > ${O}_bench_train.json 2> ${O}_bench_train.err
python - <<'PY'
import json
for f in ("bench_driver", "bench_train"):
    try:
        d = json.loads(open(f"gpurun_out/final3_r2_{f}.json").read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "value", round(d["value"] / 1e6, 2), "M/s ms/step", round(d["ms_per_step"], 5),
              "kernel_ms", r.get("kernel_ms"), "frac", r.get("frac"), "e2e", round(d["e2e"]["value"] / 1e6, 1) if d.get("e2e") else None,
              "launches", d.get("gpu_launches"))
        if "train" in d: print("   train", json.dumps(d["train"])[:330])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/final3_r2_{f}.err").read()[-1500:])
PY
exit 0
