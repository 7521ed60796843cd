#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/ab_forward_modes.py > gpurun_out/r2s_ab.json 2> gpurun_out/r2s_ab.err
cat gpurun_out/r2s_ab.json; tail -3 gpurun_out/r2s_ab.err
AB_T=50 timeout 600 python scripts/ab_forward_modes.py > gpurun_out/r2s_ab_T50.json 2>> gpurun_out/r2s_ab.err
cat gpurun_out/r2s_ab_T50.json
exit 0
