#!/bin/bash
# Round-1 third GPU pass (tight budget): new parity tests first, then the bench line, then as
# much of the full GPU suite as fits.  Usage (under gpurun): bash scripts/gpu_r3.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python -m pytest tests/test_gpu_classic_control.py tests/test_gpu_custom_env.py \
  -q --tb=short -s -p no:cacheprovider > gpurun_out/r3_new_tests.log 2>&1
echo "new tests exit $?"; tail -n 60 gpurun_out/r3_new_tests.log
timeout 60 python -m pytest tests/test_gpu_envs.py -q --tb=short -k cartpole -p no:cacheprovider 2>&1 | tail -n 5
timeout 240 python bench.py > gpurun_out/bench_r3.json 2> gpurun_out/bench_r3.err
echo "bench exit $?"; tail -c 2500 gpurun_out/bench_r3.json; tail -n 5 gpurun_out/bench_r3.err
timeout 420 python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider \
  --deselect tests/test_gpu_classic_control.py --deselect tests/test_gpu_custom_env.py \
  --durations=8 > gpurun_out/r3_full.log 2>&1
echo "full suite exit $?"; tail -n 25 gpurun_out/r3_full.log
exit 0
