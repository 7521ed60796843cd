#!/bin/bash
# pair forward + programmatic dependent launches + two-pass threshold scan: parity, A/B bench
mkdir -p gpurun_out
O=gpurun_out/r2r
timeout 1200 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_rollout.py tests/test_gpu_wide.py tests/test_gpu_envs.py tests/test_gpu_training.py -m gpu -q --tb=short -x 2>&1 | tail -12 | tee ${O}_tests.log
B="--gpus 1 --steps 20 --warmup 5 --skip-cpu-baseline --skip-ref-gpu --skip-train-probe"
WDB_OPTIONS=pdl=1 timeout 600 python bench.py $B --pair-forward > ${O}_c2_pair_pdl.json 2> ${O}_c2_pair_pdl.err
timeout 600 python bench.py $B --pair-forward > ${O}_c2_pair_nopdl.json 2> ${O}_c2_pair_nopdl.err
timeout 600 python bench.py $B > ${O}_c2_nopair.json 2> ${O}_c2_nopair.err
timeout 600 python bench.py --steps 200 --warmup 50 --skip-cpu-baseline --skip-ref-gpu --skip-train-probe > ${O}_c2_long.json 2> ${O}_c2_long.err
timeout 600 python bench.py --config 4 --steps 8 --warmup 4 --skip-cpu-baseline --skip-ref-gpu > ${O}_c4_b2.json 2> ${O}_c4_b2.err
python - <<'PY'
import json
for f in ("c2_pair_pdl", "c2_pair_nopdl", "c2_nopair", "c2_long", "c4_b2"):
    try:
        d = json.loads(open(f"gpurun_out/r2r_{f}.json").read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "value", round(d["value"] / 1e6, 2), "M/s ms/step", round(d["ms_per_step"], 5),
              "kernel_ms", r.get("kernel_ms"), "launches", d.get("gpu_launches"), "e2e", round(d["e2e"]["value"]/1e6,1), d.get("kernel_stats"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2r_{f}.err").read()[-1500:])
PY
exit 0
