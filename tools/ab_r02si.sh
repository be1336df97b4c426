#!/bin/bash
# Round 2 (second session): N = 8192 as one team per CTA, two CTAs per SM (t8192) at 1 / 4 / 8 / 16 waves of CTAs, against main.
tag=${TAG:-r02si}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
V=cognitive-radio-network_b200/variants
{
for m in welch wide; do
echo "## main 8192 $m :: $(python tools/kbench.py --nfft 8192 --mode $m --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-150)"
for g in 1 4 8 16; do
echo "## t8192 mult=$g 8192 $m :: $(CRN_GRID_MULT=$g CRN_LIB=$PWD/$V/libcrnsense_t8192.so python tools/kbench.py --nfft 8192 --mode $m --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-150)"
done; done
SIZES="2048 4096" tools/ab_all.sh noef main
} 2>&1 | tee gpurun_out/${tag}_ab.txt
