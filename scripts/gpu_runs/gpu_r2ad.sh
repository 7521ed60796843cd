#!/bin/bash
# ncu captures of the round-2 update / generic-path kernels
mkdir -p gpurun_out
O=gpurun_out/r2ad
timeout 300 ncu --set full --import-source on --clock-control none -k regex:relu_backward_bias_kernel -s 4 -c 1 \
  -o ${O}_prof_relu_bwd -f python scripts/profile_update.py > ${O}_ncu_relu.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k "regex:heads_softmax|pg_loss_kernel|pad_rows_kernel" -s 8 -c 4 \
  -o ${O}_prof_rowkernels -f python scripts/profile_update.py > ${O}_ncu_rows.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k "regex:bookkeep_kernel|gather_rows_kernel" -s 40 -c 2 \
  -o ${O}_prof_generic -f python scripts/bench_generic_path.py > ${O}_ncu_generic.log 2>&1
tail -n 1 ${O}_ncu_relu.log ${O}_ncu_rows.log ${O}_ncu_generic.log
exit 0
