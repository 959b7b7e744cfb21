#!/bin/bash
# gpurun --gpus 8: bench.py at 8 and 4 GPUs (the scaling record), NCCL_DEBUG=INFO log of rank 0 kept
mkdir -p gpurun_out/dp
for n in 8 4; do
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/dp/bench_${n}gpu.json 2> gpurun_out/dp/bench_${n}gpu.err
  echo "bench $n rc=$?"; tail -1 gpurun_out/dp/bench_${n}gpu.json | head -c 330; echo
  grep "\[0\] NCCL INFO" gpurun_out/dp/bench_${n}gpu.err | grep -i "AllReduce:.*Algo" | sort | uniq -c | sort -rn | head -8 | cut -c1-200
done
