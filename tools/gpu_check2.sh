#!/bin/bash
# One gpurun call: GPU test suite, bench (train + rollout configs[4]), timeline of the replayed graph.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/bench.json
SB_TOP=70 SB_DUMP=gpurun_out/step_kernels.txt timeout 300 python tools/graph_gaps.py > gpurun_out/graph_gaps.txt 2>&1
echo "gaps rc=$?"
head -90 gpurun_out/graph_gaps.txt
