set -x
mkdir -p gpurun_out/final
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo smoke rc=$?
timeout 400 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo bench rc=$?
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo ref rc=$?
timeout 200 python tools/ncu_step.py > gpurun_out/final/step.log 2>&1; echo step rc=$?
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final/launches.csv python tools/ncu_step.py > gpurun_out/final/ncu1.log 2>&1; echo ncu1 rc=$?
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tap_gemm2_kernel -c 4 -f -o gpurun_out/final/full_gemm2 python tools/ncu_step.py > gpurun_out/final/ncu2.log 2>&1; echo ncu2 rc=$?
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"pixel_ce_kernel|tap_wgrad_kernel" -c 4 -f -o gpurun_out/final/full_ce_wgrad python tools/ncu_step.py > gpurun_out/final/ncu3.log 2>&1; echo ncu3 rc=$?
ls -la gpurun_out/final
tail -3 gpurun_out/final/smoke.log; cat gpurun_out/final/bench.json; cat gpurun_out/final/bench_ref.json
