#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out/multi_r2_n$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > ${O}_driver_style.json 2> ${O}_driver_style.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/multi_r2_n{n}_driver_style.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"] / 1e6, 1), "M/s ms/step", round(d["ms_per_step"], 5), "reps", [round(x, 4) for x in d["ms_per_step_reps"]],
          "e2e", round(d["e2e"]["value"] / 1e6, 1), d["e2e"].get("host_cores_bound"))
    if "train" in d: print("   train", json.dumps(d["train"])[:600])
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/multi_r2_n{n}_driver_style.err").read()[-2500:])
PY
exit 0
