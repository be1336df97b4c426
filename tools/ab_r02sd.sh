#!/bin/bash
# Round 2 (second session) fourth pass: window pairs in tensor memory (main: C = 2 only; wtm8: every hybrid size; wtm0: none),
# residency of a TMEM-using kernel (ncu occupancy section at N = 2048), GPU parity suite first.
tag=${TAG:-r02sd}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "parity on variant wtm8: $(CRN_LIB=$PWD/cognitive-radio-network_b200/variants/libcrnsense_wtm8.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1)"
SIZES="${SIZES:-2048 4096 8192}" tools/ab_all.sh ${@:-notmem wtm0 main wtm8} 2>&1 | tee gpurun_out/${tag}_ab.txt
ncu --section Occupancy --section LaunchStats --metrics sm__warps_active.avg.per_cycle_active -k regex:sense_kernel -s 3 -c 1 \
  python tools/kbench.py --nfft 2048 --mode wide --steps 2 --reps 1 > gpurun_out/${tag}_occ2048.txt 2>&1
grep -i "Block Limit\|Theoretical\|Achieved\|warps_active\|Registers Per\|Shared Memory\|Waves" gpurun_out/${tag}_occ2048.txt
{
for n in 2048 4096; do for m in welch wide; do
echo "## main CRN_NO_TMA=1 $n $m :: $(CRN_NO_TMA=1 python tools/kbench.py --nfft $n --mode $m --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-150)"
done; done
for g in 2 4 8; do for m in welch wide; do
echo "## main CRN_GRID_MULT=$g 8192 $m :: $(CRN_GRID_MULT=$g python tools/kbench.py --nfft 8192 --mode $m --steps 10 --reps 3 2>&1 | tail -1 | cut -c1-150)"
done; done
} 2>&1 | tee -a gpurun_out/${tag}_ab.txt
