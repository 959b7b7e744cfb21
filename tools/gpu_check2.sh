true
for w in 1 2 3; do
  if [ $w = auto ]; then unset SB_PREP_WAVES; else export SB_PREP_WAVES=$w; fi
  timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_w$w.json 2> gpurun_out/bench.err; echo rc=$?
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_w$w.json').read().strip().splitlines()[-1])
print('waves=$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_ce']['frac'], d['rollout']['value'], d['last_loss'])
PY
done
