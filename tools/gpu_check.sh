timeout 900 python -m pytest tests -x -q -m gpu --timeout 120 --timeout-method=thread 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_step'], d['roofline_ce']['frac'], d['rollout']['value'], d['gpu_launches'], d['last_loss'])
print({k: round(v['ms'], 4) for k, v in d['roofline_by_layer'].items() if 'gate' in k})
PY
