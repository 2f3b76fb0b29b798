#!/bin/bash
# The emulated-kernel suite (tests/test_kernel_simt.py) under AddressSanitizer: the wide path's kernel SOURCE runs on the CPU (tests/hostsim/simt.h) with every
# global / "shared" memory access bounds-checked — weight rows, activation columns, outputs (numpy buffers on the ASan heap) and the static __shared__ arrays.
# No GPU needed.  Expected: all tests pass, no AddressSanitizer report.
set -e
cd "$(dirname "$0")/.."
SIM=tests/hostsim
cp $SIM/libkernsim.so /tmp/libkernsim_plain.so 2>/dev/null || true
g++ -O1 -g -std=c++20 -fPIC -shared -pthread -ffp-contract=off -w -fsanitize=address -fno-omit-frame-pointer -I/usr/local/cuda/include -x c++ $SIM/kernsim.cpp -o $SIM/libkernsim.so
touch $SIM/libkernsim.so
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python -m pytest tests/test_kernel_simt.py -x -q -p no:cacheprovider
rc=$?
[ -f /tmp/libkernsim_plain.so ] && cp /tmp/libkernsim_plain.so $SIM/libkernsim.so && touch $SIM/libkernsim.so
exit $rc
