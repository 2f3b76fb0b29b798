#!/bin/bash
# Same-box A/B of the device-resident decode leg (boxes differ by a few percent, so variants are only comparable inside one gpurun call).
# usage (under gpurun): bash tools/ab_decode.sh ["ENV=1 ENV2=2" ...]   — each argument is one variant's environment; the first run is the
# default build, and tools/ab/libb200ops_base.so (a build of an earlier commit, not tracked) is run as "base" when it exists.
B="python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 64 --warmup 8 $AB_BENCH_FLAGS"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "tok/s", round(d["ms_per_step"],4), "ms/token; matvec launches only", round(d["roofline"]["ms_per_token_matvec_only"],4), "ms")'
ROOT=$(cd "$(dirname "$0")/.." && pwd)
if [ -f "$ROOT/tools/ab/libb200ops_base.so" ]; then echo "== base"; B200_OPS_LIB=$ROOT/tools/ab/libb200ops_base.so $B 2>/dev/null | tail -1 | python -c "$P"; fi
echo "== default"; $B 2>/dev/null | tail -1 | python -c "$P"
for v in "$@"; do echo "== $v"; env $v $B 2>/dev/null | tail -1 | python -c "$P"; done
echo "== default (again)"; $B 2>/dev/null | tail -1 | python -c "$P"
