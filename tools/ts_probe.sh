#!/bin/bash
# Diagnose the --tensor-split path on a multi-GPU box: decode tok/s, graph captures / replays and scheduler splits for 1 vs 2 devices.
python oracle/make_gguf.py --config llama3-8b --ftype Q4_K_M --layers 8 --out /dev/shm/ts_probe.gguf > /dev/null
export LD_LIBRARY_PATH=oracle/_ref
for ts in "1,0" "1,1"; do
  echo "== --ts $ts"
  GGML_B200_TIMING=1 oracle/_ref/llama_drv --model /dev/shm/ts_probe.gguf --plugin llama-box_b200/libggml-b200.so --ngl 99 --ts $ts --ctx 4096 --prompt-len 512 --gen 64 --fa 2>&1 | grep -E "timing|decode_tps" | sed 's/"tokens".*//'
done
echo "== splits of one decode graph, --ts 1,1"
GGML_SCHED_DEBUG=1 LLAMA_DRV_LOG_DEBUG=1 oracle/_ref/llama_drv --model /dev/shm/ts_probe.gguf --plugin llama-box_b200/libggml-b200.so --ngl 99 --ts 1,1 --ctx 4096 --prompt-len 4 --gen 3 --fa 2>&1 | grep -E "^## SPLIT|pipeline|n_copies|graph splits" | tail -12
