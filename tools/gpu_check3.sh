timeout 300 python -m pytest tests/test_ops.py -x -q -m gpu -k frame_feedback --timeout 120 --timeout-method=thread 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['last_loss'])
PY
done
