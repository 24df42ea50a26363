timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
bash tools/profile_r02.sh r02m "c1 c3" > gpurun_out/r02m_profile.log 2>&1
for c in c1 c3; do timeout 600 python bench.py --config $c > gpurun_out/r02m_bench_$c.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02m_bench_$c.json').read().strip().splitlines()[-1])
print('$c', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1))
PY
done
du -sh gpurun_out
