T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $T bench.py --gpus 8 > gpurun_out/r02k_n8_bench_c2.json 2> gpurun_out/r02k_n8_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_n8_bench_c2.json').read().strip().splitlines()[-1])
print('N=8 value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'e2e_plugin', (d.get('e2e_plugin') or {}).get('value'))
print('per rank ms', d['timing'].get('per_rank_ms'))
print('e2e per rank', d['e2e'].get('per_rank_ms'))
print('clocks', d.get('clocks'))
PY
tail -2 gpurun_out/r02k_n8_bench_c2.err
