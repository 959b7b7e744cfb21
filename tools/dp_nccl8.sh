#!/bin/bash
# N-GPU: ms/step of the DP step under a few NCCL settings, then the timeline of the default
mkdir -p gpurun_out/dp
NP=$(nvidia-smi -L | wc -l)
run() {  # name, env...
  name=$1; shift
  env "$@" SB_NOPROF=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) tools/graph_gaps.py > gpurun_out/dp/n${NP}_$name.txt 2>&1
  echo "$name: $(grep -m1 'wall' gpurun_out/dp/n${NP}_$name.txt)"
}
run default A=1
run pool SB_DP_NCCL_POOL=1
run pool_nvls SB_DP_NCCL_POOL=1 NCCL_ALGO=NVLS
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,COLL timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29690 tools/graph_gaps.py > gpurun_out/dp/n${NP}_timeline.txt 2>&1
grep -m1 wall gpurun_out/dp/n${NP}_timeline.txt
grep -i 'nccl' gpurun_out/dp/n${NP}_timeline.txt | grep -v 'NCCL INFO' | head -12
