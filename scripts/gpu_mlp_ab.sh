#!/bin/bash
# A/B of two prebuilt libraries (libwdb200.so vs libwdb200_b.so): MLP tests, bench, launch list
TAG=$1
run() {
  timeout 300 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_rollout.py -x -q 2>&1 | tail -2
  timeout 300 python bench.py --steps 100 --warmup 20 --skip-cpu-baseline > gpurun_out/mab_${TAG}_$1.json 2> gpurun_out/mab_${TAG}_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/mab_${TAG}_$1.json').read().strip().splitlines()[-1]); print('VARIANT $1', round(d['value']/1e6,1), 'M/s step_ms', round(d['ms_per_step'],4), 'env kernel', round(d['roofline']['kernel_ms'],4))" || tail -5 gpurun_out/mab_${TAG}_$1.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mlp_forward -c 8 --csv --log-file gpurun_out/mab_${TAG}_$1.csv python bench.py --steps 3 --warmup 3 --skip-cpu-baseline --no-graph > /dev/null 2>&1
  grep mlp_forward gpurun_out/mab_${TAG}_$1.csv | awk -F'","' '{print "  mlp ns", $NF}' | head -4
}
run A
cp warp_drive_b200/libwdb200_b.so warp_drive_b200/libwdb200.so
run B
