#!/bin/bash
# One GPU-box iteration: parity tests, a short bench, ncu captures of the two hot kernels.
# Usage (under gpurun): bash scripts/gpu_round.sh <tag> [notests]
TAG=${1:-x}
mkdir -p gpurun_out
if [ "$2" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --tb=short --durations=6 2>&1 | tail -30
fi
timeout 300 python bench.py --steps 100 --warmup 20 --skip-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print("BENCH", d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), d["roofline"], d.get("kernel_stats"), d.get("gpu_launches"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$TAG.err").read()[-3000:])
PY
for CT in; do
  timeout 300 python bench.py --steps 100 --warmup 20 --skip-cpu-baseline --cta-threads $CT > gpurun_out/bench_${TAG}_ct$CT.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${TAG}_ct$CT.json').read().strip().splitlines()[-1]); print('CT$CT', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tag_continuous_kernel -s 30 -c 1 \
  -o gpurun_out/prof_fused_$TAG -f python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_fused_$TAG.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:mlp_forward_kernel -s 4 -c 2 \
  -o gpurun_out/prof_mlp_$TAG -f python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_mlp_$TAG.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --no-graph > gpurun_out/ncu_list_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_fused_$TAG.log; tail -n 3 gpurun_out/ncu_mlp_$TAG.log; exit 0
