B="python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 64 --warmup 8"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["ms_per_step"],4), "matvec-only", round(d["roofline"]["ms_per_token_matvec_only"],4))'
echo "== base";  B200_OPS_LIB=$PWD/scratch/libb200ops_base.so $B 2>/dev/null | tail -1 | python -c "$P"
echo "== new lean";   $B 2>/dev/null | tail -1 | python -c "$P"
echo "== new no-lean"; B200_MMV_NO_LEAN=1 $B 2>/dev/null | tail -1 | python -c "$P"
echo "== base";  B200_OPS_LIB=$PWD/scratch/libb200ops_base.so $B 2>/dev/null | tail -1 | python -c "$P"
echo "== new lean";   $B 2>/dev/null | tail -1 | python -c "$P"
python tools/trace_decode.py > gpurun_out/r2_trace_decode_d.txt 2> gpurun_out/trace.err
python -m pytest tests/test_gpu_executor.py tests/test_gpu_plugin.py -m gpu -q -x 2>&1 | tail -2
