#!/bin/bash
# debug: per-phase timing of the persistent decode kernel under different tuning knobs
for tune in 0; do
  echo "=== GGML_B200_MK_TUNE=$tune"
  GGML_B200_MK_TUNE=$tune timeout 300 python scratch/mk_trace.py 32 512 2>&1 | grep -E "mk trace|slot|launches"
done
