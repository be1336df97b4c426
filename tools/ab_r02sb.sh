#!/bin/bash
# Round 2 (second session) second A/B pass: computed window at C = 2 and for the one-warp-per-frame plans, bulk-copy staging at 8192 on
# top of the computed window; ncu --set full of the 8192 wideband kernel as built.
tag=${TAG:-r02sb}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for v in c2 small tma8192; do
  echo "parity on variant $v: $(CRN_LIB=$PWD/cognitive-radio-network_b200/variants/libcrnsense_$v.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1)"
done
{
SIZES="256 512 1024" tools/ab_all.sh prev main small
SIZES="2048" tools/ab_all.sh prev main c2
SIZES="8192" tools/ab_all.sh prev main tma8192
SIZES="4096" tools/ab_all.sh prev main
} 2>&1 | tee gpurun_out/${tag}_ab.txt
ncu --set full --clock-control none --import-source on -k regex:sense_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_sense_n8192 \
  python tools/kbench.py --nfft 8192 --mode wide --steps 2 --reps 1 > gpurun_out/${tag}_ncu8192.log 2>&1
tail -2 gpurun_out/${tag}_ncu8192.log
