python -m pytest tests -m gpu -q -x -k "regex or parity or split or multiline" 2>&1 | tail -3
B="python bench.py --steps 10 --no-e2e --no-cpu-baseline"
echo "== c2"; $B --config c2 2>&1 | grep -o '"kernel_ms": [0-9.]*'
echo "== c1"; $B --config c1 2>&1 | grep -o '"kernel_ms": [0-9.]*'
echo "== c1 deep8"; LC_B200_LOOKBACK_WARPS=-8 $B --config c1 2>&1 | grep -o '"kernel_ms": [0-9.]*'
echo "== c1 deep4"; LC_B200_LOOKBACK_WARPS=-4 $B --config c1 2>&1 | grep -o '"kernel_ms": [0-9.]*'
echo "== c3 deep8"; LC_B200_LOOKBACK_WARPS=-8 $B --config c3 2>&1 | grep -o '"kernel_ms": [0-9.]*'
T="python bench.py --steps 1 --warmup 3 --region-s 0.001 --no-e2e --no-cpu-baseline --config c1"
LC_B200_SPLIT_TRACE=gpurun_out/trace_c1.bin $T > /dev/null 2>&1; python tools/split_trace.py gpurun_out/trace_c1.bin
LC_B200_LOOKBACK_WARPS=-8 LC_B200_SPLIT_TRACE=gpurun_out/trace_c1_deep.bin $T > /dev/null 2>&1; python tools/split_trace.py gpurun_out/trace_c1_deep.bin
