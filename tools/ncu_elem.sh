mkdir -p gpurun_out/elem
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"reduce_bc_kernel|prep_bwd_apply_kernel|prep_fwd_kernel" -c 40 -f -o gpurun_out/elem/full_elem python tools/ncu_step.py > gpurun_out/elem/ncu.log 2>&1; echo rc=$?
ls -la gpurun_out/elem
