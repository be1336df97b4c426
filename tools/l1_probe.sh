for n in 512 1024; do
 for ex in 0 4096 8192 16384; do echo "main $n extra=$ex :: $(CRN_EXTRA_SMEM=$ex python tools/kbench.py --nfft $n --steps 10 --reps 3 | tail -1 | cut -c45-150)"; done
 echo "winl1 $n :: $(CRN_LIB=$PWD/cognitive-radio-network_b200/variants/libcrnsense_winl1.so python tools/kbench.py --nfft $n --steps 10 --reps 3 | tail -1 | cut -c45-150)"
 echo "winl1 $n wide :: $(CRN_LIB=$PWD/cognitive-radio-network_b200/variants/libcrnsense_winl1.so python tools/kbench.py --nfft $n --mode wide --steps 10 --reps 3 | tail -1 | cut -c45-150)"
 echo "main $n wide :: $(python tools/kbench.py --nfft $n --mode wide --steps 10 --reps 3 | tail -1 | cut -c45-150)"
done
for ex in 0 8192 16384; do echo "main 8192 extra=$ex :: $(CRN_EXTRA_SMEM=$ex python tools/kbench.py --nfft 8192 --steps 10 --reps 3 | tail -1 | cut -c45-150)"; done
echo "ref :: $(python tools/kbench.py --mode ref --steps 10 --reps 3 | tail -1 | cut -c45-150)"
