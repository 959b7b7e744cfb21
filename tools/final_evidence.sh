set -x
mkdir -p gpurun_out/final
timeout 600 python -m pytest tests -x -q -m gpu --timeout 120 --timeout-method=thread > gpurun_out/final/pytest_gpu.log 2>&1; echo pytest rc=$?
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo smoke rc=$?
timeout 400 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo bench rc=$?
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo ref rc=$?
timeout 200 python tools/graph_gaps.py > gpurun_out/final/graph_gaps.txt 2>&1; echo gaps rc=$?
timeout 200 python tools/ncu_step.py > gpurun_out/final/step.log 2>&1; echo step rc=$?
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final/launches.csv python tools/ncu_step.py > gpurun_out/final/ncu1.log 2>&1; echo ncu1 rc=$?
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tap_gemm2_kernel|pixel_ce_kernel|tap_wgrad_kernel" -c 8 -f -o gpurun_out/final/full_hot python tools/ncu_step.py > gpurun_out/final/ncu2.log 2>&1; echo ncu2 rc=$?
ls -la gpurun_out/final
tail -2 gpurun_out/final/pytest_gpu.log; tail -1 gpurun_out/final/smoke.log; head -c 600 gpurun_out/final/bench.json; echo; cat gpurun_out/final/bench_ref.json | head -c 400
