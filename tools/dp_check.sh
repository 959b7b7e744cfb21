#!/bin/bash
# gpurun --gpus N: DP correctness (2 ranks) then bench at every power of two up to the visible GPU count.
mkdir -p gpurun_out/dp
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_dp_nccl.py -q -m gpu -p no:cacheprovider > gpurun_out/dp/pytest_dp.log 2>&1; echo "dp test rc=$?"; tail -3 gpurun_out/dp/pytest_dp.log
for n in 8 4 2; do
  if [ $n -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/dp/bench_${n}gpu.json 2> gpurun_out/dp/bench_${n}gpu.err
    echo "bench $n rc=$?"; head -c 400 gpurun_out/dp/bench_${n}gpu.json; echo
  fi
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/graph_gaps.py > gpurun_out/dp/gaps_2gpu.txt 2>&1; echo rc=$?
grep -A40 "last step timeline" gpurun_out/dp/gaps_2gpu.txt | grep -i 'nccl\|gap\|span' | head -30
