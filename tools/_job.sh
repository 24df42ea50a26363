timeout 900 python -m pytest tests -m gpu -q -x -k "multiline or rollback or golden or c3 or remove_last" 2>&1 | tail -2
B="timeout 300 python bench.py --steps 10 --no-e2e --no-cpu-baseline"
echo "== c3"; $B --config c3 2>&1 | grep -o '"ms_per_step": [0-9.]*'
bash tools/profile_r02.sh r02l "c3" > gpurun_out/r02l_profile.log 2>&1
grep -E "ml_" gpurun_out/r02l_launches_c3.csv | tail -5 | awk -F'","' '{print substr($5,1,34), $NF}'
timeout 600 python bench.py --config c3 > gpurun_out/r02l_bench_c3.json 2>/dev/null; tail -c 300 gpurun_out/r02l_bench_c3.json
