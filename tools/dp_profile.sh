mkdir -p gpurun_out/dp
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/graph_gaps.py > gpurun_out/dp/gaps_2gpu.txt 2>&1; echo rc=$?
grep -A60 "last step timeline" gpurun_out/dp/gaps_2gpu.txt
