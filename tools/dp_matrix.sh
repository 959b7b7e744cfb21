#!/bin/bash
# 2-GPU sweep: SMs reserved for the all-reduce x NCCL CTA cap -> ms/step of the replayed DP training step
mkdir -p gpurun_out/dp
NP=${NP:-2}
port=29600
for cta in ${CTAS:-unset 8 16}; do
  for res in ${RES:-0 8 16}; do
    port=$((port+1))
    if [ $cta = unset ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$cta; fi
    SB_NOPROF=1 SB_DP_SM_RESERVE=$res timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $port tools/graph_gaps.py > gpurun_out/dp/matrix_${cta}_${res}.txt 2>&1
    echo "NCCL_MAX_CTAS=$cta reserve=$res: $(grep -m1 'wall' gpurun_out/dp/matrix_${cta}_${res}.txt)"
  done
done
