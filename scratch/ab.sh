B="python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 64 --warmup 8"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["ms_per_step"],4), "matvec-only", round(d["roofline"]["ms_per_token_matvec_only"],4))'
echo "== new";   $B 2>/dev/null | tail -1 | python -c "$P"
echo "== new, lm_head through mmvq (no persistent kernel)"; GGML_B200_DISABLE_MEGAKERNEL=1 $B 2>/dev/null | tail -1 | python -c "$P"
echo "== Q8_0 q8_0 new"; $B --ftype Q8_0 --kv q8_0 2>/dev/null | tail -1 | python -c "$P"
echo "== Q8_0 q8_0 no lean"; B200_MMV_NO_LEAN=1 $B --ftype Q8_0 --kv q8_0 2>/dev/null | tail -1 | python -c "$P"
echo "== tinyllama Q4_0 new"; $B --model tinyllama-1.1b --ftype Q4_0 --ctx 2048 2>&1 | tail -1 | python -c "$P"
echo "== tinyllama Q4_0 no lean"; B200_MMV_NO_LEAN=1 B200_FA_NO_LEAN=1 $B --model tinyllama-1.1b --ftype Q4_0 --ctx 2048 2>&1 | tail -1 | python -c "$P"
python -m pytest tests/test_gpu_executor.py tests/test_gpu_plugin.py tests/test_gpu_product.py -m gpu -q -x --deselect tests/test_gpu_product.py::test_tensor_split 2>&1 | tail -3
