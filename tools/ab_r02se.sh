#!/bin/bash
# Round 2 (second session) fifth pass: table rows fetched from TMEM one row ahead, plain loads at every hybrid size (main) against the
# previous build (wtm8) and keep-own at C = 8 on top (ko8); ncu --set full of the 8192 wideband kernel.
tag=${TAG:-r02se}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "parity on variant ko8: $(CRN_LIB=$PWD/cognitive-radio-network_b200/variants/libcrnsense_ko8.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1)"
SIZES="${SIZES:-2048 4096 8192}" tools/ab_all.sh ${@:-prev wtm8 main ko8} 2>&1 | tee gpurun_out/${tag}_ab.txt
ncu --set full --clock-control none --import-source on -k regex:sense_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_sense_n8192 \
  python tools/kbench.py --nfft 8192 --mode wide --steps 2 --reps 1 > gpurun_out/${tag}_ncu8192.log 2>&1
tail -1 gpurun_out/${tag}_ncu8192.log
