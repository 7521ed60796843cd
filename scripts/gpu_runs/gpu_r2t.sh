#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2t
timeout 1500 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_training.py tests/test_gpu_single_agent_rollout.py tests/test_gpu_update.py tests/test_gpu_core.py tests/test_gpu_custom_env.py -m gpu -q --tb=short -x 2>&1 | tail -12 | tee ${O}_tests.log
timeout 300 python scripts/bench_generic_path.py > ${O}_generic_path.json 2> ${O}_generic_path.err
cat ${O}_generic_path.json; tail -3 ${O}_generic_path.err
exit 0
