#!/bin/bash
# Round-2 closing pass on one B200: full GPU suite, every bench configuration, ncu captures of
# the driver's bench configuration, launch list.  Outputs under gpurun_out/final_r2_*.
mkdir -p gpurun_out
O=gpurun_out/${FINAL_TAG:-final_r2}
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 | tee ${O}_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke ok" | tee ${O}_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > ${O}_ref_arm.json 2> ${O}_ref_arm.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_driver.json 2> ${O}_bench_driver.err
timeout 900 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err
timeout 600 python bench.py --config 3 > ${O}_bench_config3.json 2> ${O}_bench_config3.err
timeout 900 python bench.py --config 4 --steps 8 --warmup 4 > ${O}_bench_config4_bpe2.json 2> ${O}_bench_config4_bpe2.err
timeout 900 python bench.py --config 4 --blocks-per-env 1 --steps 8 --warmup 4 --skip-cpu-baseline --skip-ref-gpu > ${O}_bench_config4_bpe1.json 2> ${O}_bench_config4_bpe1.err
timeout 600 python bench.py --mode train --steps 40 --warmup 10 > ${O}_bench_train.json 2> ${O}_bench_train.err
python - <<'PY'
import json
for f in ("ref_arm", "bench_driver", "bench_default", "bench_config3", "bench_config4_bpe2", "bench_config4_bpe1", "bench_train"):
    try:
        d = json.loads(open(f"gpurun_out/" + __import__("os").environ.get("FINAL_TAG","final_r2") + f"_{f}.json").read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "value", round(d["value"] / 1e6, 2), "M/s ms/step", round(d["ms_per_step"], 5),
              "kernel_ms", r.get("kernel_ms"), "frac", r.get("frac"), "e2e", round(d["e2e"]["value"] / 1e6, 1) if "e2e" in d else None,
              "launches", d.get("gpu_launches"))
        if "train" in d: print("   train", json.dumps(d["train"])[:400])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/" + __import__("os").environ.get("FINAL_TAG","final_r2") + f"_{f}.err").read()[-1500:])
PY
# ncu: the driver's bench configuration (--steps 20 --warmup 5), dominant kernel + MLP, then a launch list
timeout 400 ncu --set full --import-source on --clock-control none -k regex:tag_continuous_kernel -s 30 -c 1 \
  -o ${O}_prof_fused -f python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu-baseline --skip-ref-gpu --skip-train-probe > ${O}_ncu_fused.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:mlp_forward_kernel -s 4 -c 2 \
  -o ${O}_prof_mlp -f python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu-baseline --skip-ref-gpu --skip-train-probe > ${O}_ncu_mlp.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
  --log-file ${O}_launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu-baseline --skip-ref-gpu --skip-train-probe --no-graph > ${O}_ncu_list.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:tc_wide_kernel -s 6 -c 1 \
  -o ${O}_prof_wide -f python bench.py --config 4 --steps 4 --warmup 3 --skip-cpu-baseline --skip-ref-gpu > ${O}_ncu_wide.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:sa_rollout_kernel -s 3 -c 1 \
  -o ${O}_prof_sa -f python bench.py --config 3 --steps 20 --warmup 5 --skip-cpu-baseline > ${O}_ncu_sa.log 2>&1
tail -n 1 ${O}_ncu_fused.log ${O}_ncu_mlp.log ${O}_ncu_wide.log ${O}_ncu_sa.log
exit 0
