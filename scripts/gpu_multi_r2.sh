#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): NCCL test, rollout + e2e and train-mode bench at N ranks.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/multi_r2_n$N
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_nccl.py -x -q --tb=short 2>&1 | tail -5 | tee ${O}_nccl_test.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 20 > ${O}_rollout.json 2> ${O}_rollout.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --mode train --steps 40 --warmup 10 > ${O}_train.json 2> ${O}_train.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for f in ("rollout", "train"):
    try:
        d = json.loads(open(f"gpurun_out/multi_r2_n{n}_{f}.json").read().strip().splitlines()[-1])
        print(f, "N", d["n_gpus"], "value", round(d["value"] / 1e6, 1), "M/s ms/step", round(d["ms_per_step"], 5),
              "e2e", round(d["e2e"]["value"] / 1e6, 1) if "e2e" in d else None)
        if "train" in d: print("   train", json.dumps(d["train"])[:500])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/multi_r2_n{n}_{f}.err").read()[-2000:])
PY
exit 0
