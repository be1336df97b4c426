#!/bin/bash
# A/B timing of kernel variants on the GPU box: tools/ab.sh "<env assignments>" ... (each run: every FFT size, Welch K=64)
for envs in "$@"; do
  echo "#### $envs"
  for n in 512 1024 2048 4096 8192; do
    env $envs python tools/kbench.py --nfft $n --samples 1e9 --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-150
  done
  env $envs python tools/kbench.py --mode ref --samples 1e9 --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-150
  env $envs python tools/kbench.py --nfft 8192 --mode wide --samples 1e9 --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-150
done
