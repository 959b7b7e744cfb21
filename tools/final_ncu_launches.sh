#!/bin/bash
# Round-end evidence, part B: ncu launch list (duration + tensor-pipe activity per launch) of ONE eager training step.
mkdir -p gpurun_out/final
timeout 300 python tools/ncu_step.py > gpurun_out/final/step_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/final/launches.csv python tools/ncu_step.py > gpurun_out/final/ncu1.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/final/launches.csv
