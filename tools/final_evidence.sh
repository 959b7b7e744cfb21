#!/bin/bash
# Round-end evidence, part A (no profiler): GPU test suite, smoke(), bench.py (own arm + reference arm), timelines.
set -x
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/final/pytest_gpu.log 2>&1; echo pytest rc=$?
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo smoke rc=$?
timeout 600 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo ref rc=$?
timeout 600 python bench.py --agents 32 --n-action 9 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_a32_n9.json 2> gpurun_out/final/bench_a32_n9.err; echo cfg4 rc=$?
SB_TOP=80 SB_DUMP=gpurun_out/final/step_kernels.txt timeout 300 python tools/graph_gaps.py > gpurun_out/final/graph_gaps.txt 2>&1; echo gaps rc=$?
SB_DUMP=gpurun_out/final/rollout_kernels.txt timeout 300 python tools/rollout_gaps.py 16 > gpurun_out/final/rollout_gaps.txt 2>&1; echo rollout rc=$?
tail -2 gpurun_out/final/pytest_gpu.log; tail -1 gpurun_out/final/smoke.log; head -c 400 gpurun_out/final/bench.json; echo; head -c 300 gpurun_out/final/bench_ref.json; echo; head -c 300 gpurun_out/final/bench_a32_n9.json
