#!/bin/bash
# Round 2 (second session) A/B pass on one box: GPU parity suite on the in-tree library, a parity subset on the keep-own-8 variant,
# then tools/ab_all.sh over the variants given (default: the set of the computed-window experiment).
tag=${TAG:-r02sa}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for v in ${PARITY_VARIANTS:-ko8}; do
  echo "parity on variant $v: $(CRN_LIB=$PWD/cognitive-radio-network_b200/variants/libcrnsense_$v.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1)"
done
SIZES="${SIZES:-4096 8192}" tools/ab_all.sh ${@:-prev tab main calc_foldb ko8} 2>&1 | tee gpurun_out/${tag}_ab.txt
