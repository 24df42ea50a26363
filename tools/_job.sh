timeout 600 python bench.py --config c4 --no-cpu-baseline 2> gpurun_out/c4.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('c4 e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'value', d['value'])
"; tail -3 gpurun_out/c4.err
