timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
bash tools/profile_r02.sh r02k > gpurun_out/r02k_profile.log 2>&1
du -sh gpurun_out
for c in c4 c5 c1 c3 c2; do
  timeout 900 python bench.py --config $c > gpurun_out/r02k_bench_$c.json 2> gpurun_out/r02k_bench_$c.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02k_bench_$c.json').read().strip().splitlines()[-1])
    print('$c', d['metric'], round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1) if d.get('e2e') else None, 'cpu', d.get('cpu_baseline',{}).get('value'))
except Exception as e:
    print('$c failed', e); print(open('gpurun_out/r02k_bench_$c.err').read()[-1500:])
PY
done
