B="python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 64 --warmup 8"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["ms_per_step"],4), "matvec-only", round(d["roofline"]["ms_per_token_matvec_only"],4))'
echo "== new";   $B 2>/dev/null | tail -1 | python -c "$P"
echo "== new, attention not lean"; B200_FA_NO_LEAN=1 $B 2>/dev/null | tail -1 | python -c "$P"
echo "== new, min split 64"; B200_FA_MIN_SPLIT=64 $B 2>/dev/null | tail -1 | python -c "$P"
echo "== new, min split 48"; B200_FA_MIN_SPLIT=48 $B 2>/dev/null | tail -1 | python -c "$P"
echo "== new";   $B 2>/dev/null | tail -1 | python -c "$P"
echo "== new q8_0 kv"; $B --kv q8_0 2>/dev/null | tail -1 | python -c "$P"
echo "== new depth 3900"; $B --n-past 3900 2>/dev/null | tail -1 | python -c "$P"
python tools/trace_decode.py > gpurun_out/r2_trace_decode_e.txt 2> gpurun_out/trace.err
python -m pytest tests/test_gpu_ops.py tests/test_gpu_executor.py -m gpu -q -x -k "attn or attention or executor or fused" 2>&1 | tail -2
