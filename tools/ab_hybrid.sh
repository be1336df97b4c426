#!/bin/bash
# A/B of library variants on the hybrid plans: tools/ab_hybrid.sh <variant names...>  ("main" = the in-tree library)
V=cognitive-radio-network_b200/variants
for a in "--nfft 2048" "--nfft 2048 --mode wide" "--nfft 4096" "--nfft 4096 --mode wide" "--nfft 8192" "--nfft 8192 --mode wide"; do
  for v in "$@"; do
    if [ "$v" = main ]; then lib=""; else lib="CRN_LIB=$PWD/$V/libcrnsense_$v.so"; fi
    echo "## $v $a :: $(env $lib python tools/kbench.py $a --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-125)"
  done
done
