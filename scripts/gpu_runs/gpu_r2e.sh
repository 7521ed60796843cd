#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config 3 > gpurun_out/r2e_c3.json 2> gpurun_out/r2e_c3.err
timeout 600 python bench.py --config 3 --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2e_c3_drv.json 2> gpurun_out/r2e_c3_drv.err
timeout 600 python bench.py --mode train --steps 40 --warmup 10 > gpurun_out/r2e_train1.json 2> gpurun_out/r2e_train1.err
python - <<'PY'
import json
for f in ("r2e_c3", "r2e_c3_drv", "r2e_train1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]/1e6, 1), "M/s ms/step", round(d["ms_per_step"], 5), "launches", d.get("gpu_launches"))
        for k in ("roofline", "train", "ref_gpu", "cpu_baseline", "e2e"):
            if k in d: print("   ", k, json.dumps(d[k])[:700])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
exit 0
