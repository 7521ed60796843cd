#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_gpu_training.py tests/test_gpu_nccl.py -x -q --tb=short 2>&1 | tail -15
timeout 600 python bench.py --mode train --steps 40 --warmup 10 > gpurun_out/r2f_train1.json 2> gpurun_out/r2f_train1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --mode train --steps 40 --warmup 10 > gpurun_out/r2f_train2.json 2> gpurun_out/r2f_train2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/r2f_roll2.json 2> gpurun_out/r2f_roll2.err
python - <<'PY'
import json
for f in ("r2f_train1", "r2f_train2", "r2f_roll2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]/1e6, 1), "M/s ms/step", round(d["ms_per_step"], 5))
        for k in ("train", "e2e"):
            if k in d: print("   ", k, json.dumps(d[k])[:600])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
exit 0
