#!/bin/bash
# debug: per-step logits deviation CPU vs plugin for pure decode
set -e
export LD_LIBRARY_PATH=$PWD/oracle/_ref
FT=${1:-Q8_0}; KV=${2:-q8_0}; N=${3:-16}
python oracle/make_gguf.py --config test-small --ftype $FT --weights gauss --out /dev/shm/m.gguf >/dev/null
oracle/_ref/llama_drv --model /dev/shm/m.gguf --ctx 512 --prompt-len 1 --gen $N --logits-out /dev/shm/c.bin --fa --ctk $KV --ctv $KV --ngl 0 --threads 16 --no-repack --dump /dev/shm/c.dump
GGML_BACKEND_PATH=$PWD/llama-box_b200/libggml-b200.so oracle/_ref/llama_drv --model /dev/shm/m.gguf --ctx 512 --prompt-len 1 --gen $N --logits-out /dev/shm/g.bin --fa --ctk $KV --ctv $KV --plugin $PWD/llama-box_b200/libggml-b200.so --ngl 99 --dump /dev/shm/g.dump
python scratch/cmp_logits.py /dev/shm/g.bin /dev/shm/c.bin $N
python - <<'PY'
a=open('/dev/shm/c.dump').read().splitlines(); b=open('/dev/shm/g.dump').read().splitlines()
print(len(a),len(b))
import re
def key(l): return l.split('sum=')[0]
bi=0; shown=0
from collections import defaultdict
# align by name sequence
bm=defaultdict(list)
for l in b: bm[l.split()[0]].append(l)
cnt=defaultdict(int)
for l in a:
    n=l.split()[0]; k=cnt[n]; cnt[n]+=1
    if k < len(bm[n]):
        m=bm[n][k]
        sa=float(re.search(r'abs=(\S+)',l).group(1)); sb=float(re.search(r'abs=(\S+)',m).group(1))
        if abs(sa-sb) > 1e-5*abs(sa) and shown<25:
            print('C',l); print('G',m); shown+=1
PY
