#!/bin/bash
# A/B of library variants on every FFT size: tools/ab_all.sh <variant names...>  ("main" = the in-tree library)
# Sizes/modes can be narrowed with SIZES="2048 8192" MODES="welch wide".
V=cognitive-radio-network_b200/variants
for n in ${SIZES:-256 512 1024 2048 4096 8192}; do
  for mode in ${MODES:-welch wide}; do
    for v in "$@"; do
      if [ "$v" = main ]; then lib=""; else lib="CRN_LIB=$PWD/$V/libcrnsense_$v.so"; fi
      echo "## $v $n $mode :: $(env $lib python tools/kbench.py --nfft $n --mode $mode --steps ${STEPS:-10} --reps 3 2>&1 | tail -1 | cut -c1-150)"
    done
  done
done
