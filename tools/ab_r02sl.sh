#!/bin/bash
# Round 2 (second session): TMEM tables / window / accumulators for the one-warp-per-frame plans (variant small256)
tag=${TAG:-r02sl}
mkdir -p gpurun_out
V=$PWD/cognitive-radio-network_b200/variants
echo "parity on variant small256: $(CRN_LIB=$V/libcrnsense_small256.so timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -1)"
{
SIZES="256 512 1024" tools/ab_all.sh main small256
for v in main small256; do
  if [ "$v" = main ]; then lib=""; else lib="CRN_LIB=$V/libcrnsense_$v.so"; fi
  echo "## $v 512 ref :: $(env $lib python tools/kbench.py --mode ref --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-150)"
  echo "## $v 1024 welch allbins :: $(env $lib CRN_NO_PRUNE=1 python tools/kbench.py --nfft 1024 --mode welch --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-150)"
done
} 2>&1 | tee gpurun_out/${tag}_ab.txt
