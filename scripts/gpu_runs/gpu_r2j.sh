#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py -x -q --tb=short -k cartpole_learns 2>&1 | tail -8
timeout 600 python scripts/bench_small_kernels.py > gpurun_out/r2j_small_kernels.json 2> gpurun_out/r2j_small_kernels.err; tail -c 1500 gpurun_out/r2j_small_kernels.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tc_wide_kernel -s 6 -c 1 \
  -o gpurun_out/prof_wide_r2j -f python bench.py --config 4 --steps 4 --warmup 3 --skip-cpu-baseline --skip-ref-gpu > gpurun_out/ncu_wide_r2j.log 2>&1
tail -n 2 gpurun_out/ncu_wide_r2j.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sa_rollout_kernel -s 3 -c 1 \
  -o gpurun_out/prof_sa_r2j -f python bench.py --config 3 --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/ncu_sa_r2j.log 2>&1
tail -n 2 gpurun_out/ncu_sa_r2j.log
exit 0
