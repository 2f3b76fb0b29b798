#!/bin/bash
# debug: plugin decode with / without the persistent kernel vs ggml-cpu
export LD_LIBRARY_PATH=$PWD/oracle/_ref
FT=${1:-Q4_0}; KV=${2:-f16}; N=${3:-4}
python oracle/make_gguf.py --config test-small --ftype $FT --weights gauss --out /dev/shm/m.gguf >/dev/null
oracle/_ref/llama_drv --model /dev/shm/m.gguf --ctx 512 --prompt-len 1 --gen $N --logits-out /dev/shm/c.bin --fa --ctk $KV --ctv $KV --ngl 0 --threads 16 --no-repack 2>/dev/null | cut -c1-200
P=$PWD/llama-box_b200/libggml-b200.so
GGML_BACKEND_PATH=$P oracle/_ref/llama_drv --model /dev/shm/m.gguf --ctx 512 --prompt-len 1 --gen $N --logits-out /dev/shm/g.bin --fa --ctk $KV --ctv $KV --plugin $P --ngl 99 2>/dev/null | cut -c1-200
GGML_B200_DISABLE_MEGAKERNEL=1 GGML_BACKEND_PATH=$P oracle/_ref/llama_drv --model /dev/shm/m.gguf --ctx 512 --prompt-len 1 --gen $N --logits-out /dev/shm/h.bin --fa --ctk $KV --ctv $KV --plugin $P --ngl 99 2>/dev/null | cut -c1-200
echo "mega vs cpu"; python scratch/cmp_logits.py /dev/shm/g.bin /dev/shm/c.bin $N
echo "per-op vs cpu"; python scratch/cmp_logits.py /dev/shm/h.bin /dev/shm/c.bin $N
