#!/bin/bash
# compute-sanitizer passes over small instances of every kernel family (run on the GPU box).
set -u
for tool in memcheck racecheck; do
  for args in "--nfft 1024 --samples 1e6" "--mode ref --samples 5e5" "--nfft 512 --navg 7 --samples 5e5" "--nfft 8192 --mode wide --samples 2e6" "--nfft 2048 --navg 10 --samples 1e6" "--nfft 2048 --navg 64 --samples 3e6" "--nfft 2048 --mode wide --navg 8 --samples 1e6 --window rect" "--nfft 4096 --samples 2e6" "--nfft 4096 --mode wide --samples 2e6" "--nfft 8192 --samples 2e6" "--nfft 256 --mode wide --samples 5e5"; do
    echo "== $tool $args"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/kbench.py $args --steps 1 --reps 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|sense_" | head -6
  done
done
