#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2z
timeout 900 python -m pytest tests/test_gpu_envs.py tests/test_gpu_rollout.py -m gpu -q --tb=short -x 2>&1 | tail -4 | tee ${O}_tests.log
AB_T=50 AB_MODES=fork+tail,fork timeout 600 python scripts/ab_forward_modes.py > ${O}_ab_tail.json 2> ${O}_ab.err
cat ${O}_ab_tail.json; tail -2 ${O}_ab.err
exit 0
