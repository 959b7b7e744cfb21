#!/bin/bash
# Round-end evidence, part C: ncu --set full of the hot kernels of one eager training step (a few launches of each).
mkdir -p gpurun_out/final
timeout 300 python tools/ncu_step.py > gpurun_out/final/step_plain2.log 2>&1; echo "plain rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"^tap_gemm2_kernel|^tap_wgrad_kernel|^dense_kernel|^prep_fwd_kernel" -c 24 -f -o gpurun_out/final/full_hot python tools/ncu_step.py > gpurun_out/final/ncu2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/final/full_hot.ncu-rep
