timeout 900 python -m pytest tests -m gpu -q -x -k "multiline or rollback or golden or c3" 2>&1 | tail -2
B="timeout 300 python bench.py --steps 10 --no-e2e --no-cpu-baseline"
echo "== c3"; $B --config c3 2>&1 | grep -o '"ms_per_step": [0-9.]*'
N="ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 100 --csv"
B2="python bench.py --steps 2 --warmup 3 --region-s 0.01 --no-e2e --no-cpu-baseline"
$N --log-file gpurun_out/r02h_launches_c3_warm.csv $B2 --config c3 > /dev/null 2>&1; grep -E "ml_" gpurun_out/r02h_launches_c3_warm.csv | tail -5 | awk -F'","' '{print substr($5,1,34), $NF}'
