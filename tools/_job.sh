timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "chain" 2>&1 | tail -3
timeout 600 python bench.py --config c4 --no-cpu-baseline 2> gpurun_out/c4.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('c4 e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'value', d['value'], 'stale', d['roofline']['traffic_source'])
"; tail -2 gpurun_out/c4.err
